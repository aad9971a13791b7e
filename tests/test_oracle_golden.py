"""The oracle against every known answer the reference's own tests hold for the bzip2 path
(SURVEY.md section 8c), the README sizes (legacy-V8 sort) and the survey's provisional goldens."""
import bz2
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import HERE, fixture_bytes

# NPM/test/bwtest.js:39-79
BWT_VECTORS = [
    ("bcababa", "cbbaaab", 5),
    ("ABCDEFGHIJKLMNOPQRSTUVWXYZ", "ZABCDEFGHIJKLMNOPQRSTUVWXY", 0),
    ("ZYXWVUTSRQPONMLKJIHGFEDCBA", "BCDEFGHIJKLMNOPQRSTUVWXYZA", 25),
    ("SIX.MIXED.PIXIES.SIFT.SIXTY.PIXIE.DUST.BOXES", "TEXYDST.E.IXIXIXXSSMPPS.B..E.S.EUSFXDIIOIIIT", 29),
    ("Mary had a little lamb, its fleece was white as snow" * 8 + "Nary had a little lamb, its fleece was white as snow",
     "dddddddddeeeeeeeeesssssssssyyyyyyyyy,,,,,,,,,eeeeeeeeeaaaaaaaaassssssssseeeeeeeeesss"
     "ssssssbbbbbbbbbwwwwwwwww         hhhhhhhhhlllllllllNMMMMMMMM         wwwwwwwwwmmmmmm"
     "mmmeeeeeeeeeaaaaaaaaatttttttttlllllllllccccccccceeeeeeeeelllllllll                  "
     "wwwwwwwwwhhhhhhhhh         lllllllll         tttttttttfffffffff         aaaaaaaaasss"
     "ssssssnnnnnnnnnaaaaaaaaatttttttttaaaaaaaaaaaaaaaaaa         iiiiiiiiitttttttttiiiiii"
     "iiiiiiiiiiiiooooooooo                  rrrrrrrrr", 99),
]
FIB = [0, 1, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765, 10946, 17711, 28657, 46368, 75025,
       121393, 196418, 317811, 514229, 832040, 1346269, 2178309, 3524578, 5702887, 9227465, 14930352]


@pytest.mark.parametrize("inp,out,idx", BWT_VECTORS)
def test_bwt_known_answers(oracle, inp, out, idx):
    assert oracle.bwt(inp.encode()) == (out.encode(), idx)


def test_bwt_tie_rule(oracle):
    assert oracle.bwt(b"aaaa") == (b"aaaa", 3)
    assert oracle.bwt(b"abab") == (b"bbaa", 1)
    assert oracle.bwt(b"abcabcabc") == (b"cccaaabbb", 2)
    assert oracle.bwt(b"x") == (b"x", 0)


def test_bwt_matches_bruteforce(oracle):
    rng = np.random.default_rng(5)
    for n in [2, 3, 5, 8, 17, 64, 200]:
        for nsym in (1, 2, 3, 256):
            t = rng.integers(0, nsym, n, dtype=np.uint8).tobytes()
            rots = sorted(range(n), key=lambda i: (t[i:] + t[:i], -i))
            exp = bytes(t[i - 1] for i in rots)
            assert oracle.bwt(t) == (exp, rots.index(0))


HUFF_VECTORS = [  # NPM/test/huffman.js:16-76: (frequencies ascending, length limit, expected code lengths)
    ([1], 32, [1]),
    ([1, 1], 32, [1, 1]),
    ([1] * 5, 32, [3, 3, 2, 2, 2]),
    ([0, 0, 1, 1, 1, 1], 3, [3, 3, 3, 3, 2, 2]),
    (FIB[:36], 20, [20] * 16 + [19, 19, 18, 17, 16, 16, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]),
    (FIB[:22], 20, [20, 20, 19, 19, 19, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]),
    (FIB[:21], 20, [20, 20, 19, 18, 17, 16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1]),
    (FIB[:36], 6, [6] * 30 + [5, 5, 5, 4, 3, 2]),
]


def test_huffman_allocator_known_answers(oracle):  # NPM/test/huffman.js:16-76
    for freqs, limit, expect in HUFF_VECTORS:
        assert oracle.huff_alloc(freqs, limit) == expect


def test_fls(oracle):  # NPM/test/test-fls.js:14-47
    assert [oracle.fls(v) for v in (0, 1, 2, 3, 4)] == [0, 1, 2, 2, 3]
    for v in list(range(258)) + [(1 << i) + d for i in range(1, 31) for d in (-2, -1, 0, 1, 2)]:
        assert oracle.fls(v) == v.bit_length()
    assert oracle.fls(0x7FFFFFFF) == 31 and oracle.fls(0xFFFFFFFF) == 32 and oracle.fls(0x1FFFFFFFF) == 33
    assert oracle.fls(0x1FFFFFFFFFFFFF) == 53 and oracle.fls(0x20000000000000) == 54


@pytest.mark.parametrize("n", range(5))
def test_decode_fixtures(oracle, n):  # NPM/test/bzip2-basic.js
    assert oracle.decompress(fixture_bytes(f"sample{n}.bz2")) == fixture_bytes(f"sample{n}.ref")


@pytest.mark.parametrize("n", range(5))
def test_table_fixtures(oracle, n):  # NPM/test/bzip2-table.js
    t = oracle.table(fixture_bytes(f"sample{n}.bz2"))
    assert "".join(f"{p}\t{s}\n" for p, s in t) == fixture_bytes(f"sample{n}.bzt").decode()


def test_block_fixtures(oracle):  # NPM/test/bzip2-block.js
    assert oracle.decompress_block(fixture_bytes("sample0.bz2"), 32) == b"This is a test\n"
    for f, b in [("sample2", 544888), ("sample4", 32), ("sample4", 1596228), ("sample4", 2342106)]:
        assert oracle.decompress_block(fixture_bytes(f + ".bz2"), b) == fixture_bytes(f"{f}.{b}")


def test_readme_sizes_legacy_v8(oracle):
    """/root/reference/README.md:112,115 -- the only published known answers for the compressor."""
    s5 = fixture_bytes("sample5.ref")
    assert len(oracle.compress(s5, 9, oracle.SORT_LEGACY_V8, threads=4)) == 275087
    assert len(oracle.compress(s5, 1, oracle.SORT_LEGACY_V8, threads=4)) == 341615


SURVEY_GOLDENS = {  # SURVEY.md appendix C (stable sort): (size, sha256[:32])
    ("sample0", 1): (57, "d2cf1f9848c5a87cc0299157f83cba16"), ("sample0", 9): (57, "1aa93d50340ba8253c826cfee6ab4a44"),
    ("sample1", 9): (32860, "90a6638f1aa9d84843f94877b0b5af41"), ("sample2", 1): (80144, "b380a4a823a0fee2e12d06aebd6bdfaf"),
    ("sample2", 2): (74702, "e82a5253588be53cfbc87571f5cd6c78"), ("sample3", 1): (275, "5390f4b1097462f9a4c26b191d7b8c13"),
    ("sample3", 9): (235, "cacd0a28bb432289d1d3a0d185e31e55"), ("sample4", 1): (304142, "19be4a72b305f10074836918ebdd0fa5"),
    ("sample4", 9): (334734, "dd429c976404f1c6e895f68e9f2cde58"), ("sample5", 1): (341540, "b7f354a278dc141346a579723a5db2d2"),
    ("sample5", 5): (292452, "09c8e4ae0e121836d55cce558936172d"), ("sample5", 9): (274768, "236be53bab8972f04032ef8851adad95"),
}


@pytest.mark.parametrize("key", sorted(SURVEY_GOLDENS))
def test_survey_goldens_and_roundtrip(oracle, key):
    name, level = key
    data = fixture_bytes(name + ".ref")
    comp = oracle.compress(data, level, threads=4)
    size, sha = SURVEY_GOLDENS[key]
    assert (len(comp), hashlib.sha256(comp).hexdigest()[:32]) == (size, sha)
    assert bz2.decompress(comp) == data          # libbz2 accepts the stream
    assert oracle.decompress(comp) == data       # and so does the restated reference decoder


TINY = {  # SURVEY.md appendix C, whole output in hex
    b"": "425a683917724538509000000000",
    b"Q": "425a6839314159265359cda1f6fb00000002002000200021184682ee48a70a1219b43edf60",
    b"aaaa": "425a6839314159265359881233a600000241004000200020002127a820538bb9229c2848440919d300",
    b"aaaaa": "425a683931415926535944a4303d00000241002000200020002127a820538bb9229c28482252181e80",
    b"a" * 256: "425a6839314159265359efac2e370000008100a0000008200021008293177245385090efac2e37",
    b"a" * 255 + b"x": "425a6839314159265359817405480000000180a000004000082000219ea0660b9c5dc914e1424205d01520",
    bytes(1000): "425a683931415926535901e8d060000001c001c00000800008200020aa3d0662ea0c2ee48a70a12003d1a0c0",
}


def test_tiny_streams(oracle):
    for data, hx in TINY.items():
        assert oracle.compress(data, 9).hex() == hx
        assert oracle.decompress(bytes.fromhex(hx)) == data
    assert len(oracle.compress(bytes(2_000_000), 9)) == 49
    assert len(oracle.compress(b"ab" * 500_000, 9)) == 73
    assert len(oracle.compress(b"abc" * 333_333, 9)) == 84


def test_errors_and_level(oracle):
    with pytest.raises(oracle.OracleError) as e:
        oracle.compress(b"x", 0)
    assert e.value.errorCode == -100
    for bad in (b"", b"BZ", b"BZh0" + bytes(20), b"hello world!"):
        with pytest.raises(oracle.OracleError) as e:
            oracle.decompress(bad)
        assert e.value.errorCode == -2
    c = bytearray(oracle.compress(b"hello world, hello world"))
    c[20] ^= 0x10
    with pytest.raises(oracle.OracleError) as e:
        oracle.decompress(bytes(c))
    assert e.value.errorCode == -5


def test_mt_equals_single_thread(oracle):
    data = fixture_bytes("sample5.ref")[:700_000]
    assert oracle.compress(data, 1, threads=1) == oracle.compress(data, 1, threads=5)
    comp = oracle.compress(data, 1)
    assert oracle.decompress(comp, threads=4) == data


def test_corpus_generator_is_deterministic():
    from compressjs_flattened_b200.corpus import gen_html, gen_text
    gold = json.load(open(os.path.join(HERE, "golden", "corpus_goldens.json")))
    assert hashlib.sha256(gen_html(2_130_640, 5).tobytes()).hexdigest() == gold["html:2130640:5:L9"]["input_sha256"]
    t = gen_text(3_000_000, 8)
    assert hashlib.sha256(gen_text(10_000_000, 8)[:3_000_000].tobytes()).hexdigest() == hashlib.sha256(t.tobytes()).hexdigest()
    assert np.array_equal(gen_text(1_000_000, 8, first_chunk=2), t[2_000_000:])
