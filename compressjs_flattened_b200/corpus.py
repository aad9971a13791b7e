"""Deterministic synthetic corpora for the parity tests and bench.py.

gen_text(n, seed)  "enwik8-like": Zipf(1.05) words from a 65 536-word synthetic vocabulary
                   (letters by English frequency), wiki markup, punctuation/newlines, digit
                   groups and a <page>...<text xml:space="preserve"> boilerplate every ~6 KB.
gen_html(n, seed)  "sample5-like": Parsoid-style HTML: prose with wiki links (target spelled three times),
                   bold spans, paragraphs and citation marks carrying growing tsr/dsr source offsets,
                   runs of 4-8 spaces; ratio 0.147 at level 9 (the reference's sample5.ref: 0.129).

Chunk c (1 000 000 bytes) depends only on splitmix64(seed, c), so chunks can be produced in any
order / in parallel and an 8 GB corpus never needs a serial pass.  Pure numpy integer arithmetic:
the bytes are identical on every machine.
"""
import numpy as np

CHUNK = 1_000_000
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return x ^ (x >> np.uint64(31))


def _stream(seed, chunk, lane, count):
    """count pseudo-random uint64 for (seed, chunk, lane)."""
    base = _mix(np.array([(seed * 0x100000001B3 + chunk * 0x9E3779B1 + lane * 0x85EBCA77) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
    with np.errstate(over="ignore"):
        return _mix(base + np.arange(count, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95))


_LETTERS = np.frombuffer(
    (b"e" * 31 + b"t" * 23 + b"a" * 21 + b"o" * 19 + b"i" * 18 + b"n" * 17 + b"s" * 16 + b"h" * 15 + b"r" * 15 + b"d" * 11 +
     b"l" * 10 + b"c" * 7 + b"u" * 7 + b"m" * 6 + b"w" * 6 + b"f" * 6 + b"g" * 5 + b"y" * 5 + b"p" * 5 + b"b" * 4 + b"v" * 3 +
     b"k" * 2 + b"j" + b"x" + b"q" + b"z")[:256].ljust(256, b"e"), dtype=np.uint8)

_TABLE = None
_MAXLEN = 48


def _table():
    """String table: 65 536 vocabulary words, then numbers 0..9999, then markup pieces."""
    global _TABLE
    if _TABLE is not None:
        return _TABLE
    V = 65536
    with np.errstate(over="ignore"):
        h = _mix(np.arange(V, dtype=np.uint64) * np.uint64(0x2545F4914F6CDD1D) + np.uint64(12345))
    rank = np.arange(V)
    length = 1 + (h % np.uint64(4)).astype(np.int64) + np.minimum(9, (np.log2(rank + 2) * 0.62).astype(np.int64))
    length = np.clip(length, 1, 14)
    tab = np.zeros((V + 10000 + 96, _MAXLEN), dtype=np.uint8)
    lens = np.zeros(V + 10000 + 96, dtype=np.int64)
    for k in range(14):
        with np.errstate(over="ignore"):
            hk = _mix(h + np.uint64(k * 7919 + 1))
        tab[:V, k] = _LETTERS[(hk & np.uint64(255)).astype(np.int64)]
    lens[:V] = length
    for w, s in enumerate([b"the", b"of", b"and", b"in", b"a", b"to", b"is", b"was", b"for", b"as", b"by", b"with", b"on", b"that", b"s"]):
        tab[w, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        lens[w] = len(s)
    for v in range(10000):
        s = str(v).encode()
        tab[V + v, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        lens[V + v] = len(s)
    pieces = [b" ", b", ", b". ", b"\n", b"\n\n", b"; ", b": ", b" (", b") ", b"[[", b"]] ", b"|", b"''", b"'' ", b"'''", b"''' ",
              b"==", b"==\n", b"{{", b"}} ", b"&quot;", b"&quot; ", b"&amp; ", b"* ", b"# ", b" - ", b"</text>\n    </revision>\n  </page>\n",
              b"  <page>\n    <title>", b"</title>\n    <id>", b"</id>\n    <revision>\n      <id>", b"</id>\n      <timestamp>200",
              b"-0", b"T1", b":5", b"Z</timestamp>\n      <contributor>\n", b"        <username>", b"</username>\n        <id>",
              b"</id>\n      </contributor>\n      <text ", b"xml:space=\"preserve\">", b"    ", b"     ", b"      ", b"       ", b"        ",
              b"<div class=\"", b"<span", b"<p", b"</p>\n", b"</div>\n", b"</span>", b"<a href=\"./", b"\"", b"</a> ", b"<li", b"</li>\n",
              b" data-parsoid='{\"dsr\":[", b",", b"]}'", b"<td", b"</td>", b"<tr", b"</tr>\n", b">",
              # appended for gen_html (indices above are unchanged, so gen_text's bytes are too)
              b"<a rel=\"mw:WikiLink\" href=\"./", b"\" data-parsoid='{\"tsr\":[", b"],\"a\":{\"href\":\"./", b"\"},\"sa\":{\"href\":\"",
              b"\"},\"stx\":\"piped\",\"dsr\":[", b",2]}'>", b"</a>", b"<b data-parsoid='{\"tsr\":[", b"],\"dsr\":[", b",3,3]}'>", b"</b>",
              b"</p>\n\n<p data-parsoid='{\"dsr\":[", b",0,0]}'>", b"<span class=\"reference\" data-parsoid='{\"dsr\":[",
              b"<a href=\"#cite_note-", b"\">[", b"]</a></span>", b"_"]
    assert len(pieces) <= 96 and max(len(x) for x in pieces) <= _MAXLEN
    for i, s in enumerate(pieces):
        tab[V + 10000 + i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
        lens[V + 10000 + i] = len(s)
    w = 1.0 / np.power(np.arange(1, V + 1, dtype=np.float64), 1.05)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    _TABLE = dict(tab=tab, lens=lens, cdf=cdf, V=V, NUM=V, PIECE=V + 10000, names={s: V + 10000 + i for i, s in enumerate(pieces)})
    return _TABLE


def _assemble(ids, T, nbytes, cap=None):
    """Concatenate table strings ids[] and cut/pad to nbytes; cap[] marks ids whose first letter is upper-cased."""
    lens = T["lens"][ids]
    ends = np.cumsum(lens)
    total = int(ends[-1])
    starts = ends - lens
    tok = np.repeat(np.arange(ids.size), lens)
    off = np.arange(total) - np.repeat(starts, lens)
    out = T["tab"][ids[tok], off]
    if cap is not None:
        at = starts[cap & (ids < T["V"])]
        out[at] -= np.where((out[at] >= 97) & (out[at] <= 122), 32, 0).astype(np.uint8)
    if total < nbytes:
        out = np.concatenate([out, np.full(nbytes - total, 32, dtype=np.uint8)])
    return out[:nbytes]


def _words(u, T):
    return np.searchsorted(T["cdf"], (u >> np.uint64(11)).astype(np.float64) / float(1 << 53)).astype(np.int64).clip(0, T["V"] - 1)


def _text_chunk(seed, c):
    T = _table()
    N = T["names"]
    K = 215_000
    w = _words(_stream(seed, c, 0, K), T)
    # phrase structure: with probability ~0.55 a word is a fixed successor of the word drawn before it
    hs = _stream(seed, c, 5, K)
    follow = (hs % np.uint64(100)) < np.uint64(55)
    follow[0] = False
    with np.errstate(over="ignore"):
        succ = _words(_mix(np.roll(w, 1).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + ((hs >> np.uint64(40)) % np.uint64(3))), T)
    w = np.where(follow, succ, w)
    r = (_stream(seed, c, 1, K) % np.uint64(1000)).astype(np.int64)
    # separator after each word
    sep = np.full(K, N[b" "], dtype=np.int64)
    for lo, hi, s in [(0, 38, b", "), (38, 70, b". "), (70, 84, b"\n"), (84, 90, b"\n\n"), (90, 96, b"; "), (96, 101, b": "),
                      (101, 107, b" ("), (107, 113, b") "), (113, 118, b" - "), (118, 121, b"&amp; ")]:
        sep[(r >= lo) & (r < hi)] = N[s]
    # optional markup wrapped around the word
    r2 = (_stream(seed, c, 2, K) % np.uint64(1000)).astype(np.int64)
    pre = np.full(K, -1, dtype=np.int64)
    post = np.full(K, -1, dtype=np.int64)
    for lo, hi, a, b in [(0, 30, b"[[", b"]] "), (30, 38, b"''", b"'' "), (38, 42, b"'''", b"''' "), (42, 47, b"==", b"==\n"),
                         (47, 53, b"{{", b"}} "), (53, 58, b"&quot;", b"&quot; ")]:
        m = (r2 >= lo) & (r2 < hi)
        pre[m] = N[a]
        post[m] = N[b]
    # 1.5 % of the words are digit groups
    num = (r2 >= 985)
    w = np.where(num, T["NUM"] + (_stream(seed, c, 3, K) % np.uint64(10000)).astype(np.int64), w)
    seq = np.stack([pre, w, np.where(post >= 0, post, sep)], axis=1)
    # page boilerplate every ~1100 words (~6 KB)
    hb = _stream(seed, c, 4, K // 1100 + 2)
    rows = []
    for j in range(K // 1100 + 1):
        h = int(hb[j])
        d = lambda k: T["NUM"] + ((h >> (k * 9)) % 10000)
        rows.append((j * 1100, [N[b"</text>\n    </revision>\n  </page>\n"], N[b"  <page>\n    <title>"], int(w[j * 1100]), N[b" "], int(w[j * 1100 + 1]),
                                N[b"</title>\n    <id>"], d(0), N[b"</id>\n    <revision>\n      <id>"], d(1), d(2),
                                N[b"</id>\n      <timestamp>200"], T["NUM"] + (h % 7), N[b"-0"], T["NUM"] + 1 + (h >> 7) % 9, N[b"-0"], T["NUM"] + 1 + (h >> 11) % 9,
                                N[b"T1"], T["NUM"] + (h >> 15) % 10, N[b":5"], T["NUM"] + (h >> 19) % 10, N[b":5"], T["NUM"] + (h >> 23) % 10,
                                N[b"Z</timestamp>\n      <contributor>\n"], N[b"        <username>"], int(w[j * 1100 + 2]),
                                N[b"</username>\n        <id>"], d(3), N[b"</id>\n      </contributor>\n      <text "], N[b"xml:space=\"preserve\">"]]))
    flat = seq.reshape(-1)
    parts, last = [], 0
    for pos, ids in rows:
        parts.append(flat[last * 3:pos * 3])
        parts.append(np.array(ids, dtype=np.int64))
        last = pos
    parts.append(flat[last * 3:])
    ids = np.concatenate(parts)
    ids = ids[ids >= 0]
    with np.errstate(over="ignore"):
        cap = (_mix(np.arange(ids.size, dtype=np.uint64) + np.uint64(seed * 977 + c)) % np.uint64(100)) < np.uint64(9)
    return _assemble(ids, T, CHUNK, cap)


def _html_chunk(seed, c):
    """Parsoid-like HTML: running prose (phrase structure as in gen_text, a 2 048-word vocabulary) with wiki links whose
    target is spelled three times (href, a.href, sa.href), bold spans, paragraphs and citation marks, each carrying
    data-parsoid source offsets (tsr/dsr) that grow with the text like a real position counter, plus the odd run of
    4-8 spaces.  Tuned to the ratio of the reference's sample5.ref (2 130 640 B -> 0.129 at level 9; SURVEY C1b band
    0.12-0.16)."""
    T = _table()
    N = T["names"]
    K = 90_000
    VH = 2048
    w = _words(_stream(seed, c, 10, K), T) % VH
    hs = _stream(seed, c, 13, K)
    follow = (hs % np.uint64(100)) < np.uint64(78)
    follow[0] = False
    with np.errstate(over="ignore"):
        succ = _words(_mix(np.roll(w, 1).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15) + ((hs >> np.uint64(40)) % np.uint64(1))), T) % VH
    w = np.where(follow, succ, w)
    r = (_stream(seed, c, 11, K) % np.uint64(1000)).astype(np.int64)
    h = _stream(seed, c, 12, K)
    wl = T["lens"][w]
    # source offsets: a counter over the wikitext this would have been rendered from
    pos = 800 + np.cumsum(wl + 1 + np.where(r < 70, 4, 0))
    end = pos + wl + np.where(r < 70, 4 + (h % np.uint64(3)).astype(np.int64) * wl, 0)

    def digits(v):  # decimal digits of v (>= 100) as table ids: leading part, tens, units
        return T["NUM"] + (v // 100) % 10000, T["NUM"] + (v // 10) % 10, T["NUM"] + v % 10

    p0, p1, p2 = digits(pos)
    e0, e1, e2 = digits(end)
    NC = 26
    cols = [np.full(K, -1, dtype=np.int64) for _ in range(NC)]
    cols[0][:] = w  # plain token: word + separator
    cols[1][:] = N[b" "]
    cols[1][(r >= 900) & (r < 945)] = N[b", "]
    cols[1][(r >= 945) & (r < 985)] = N[b". "]
    sp = r >= 996  # runs of 4-8 spaces
    cols[1][sp] = N[b"    "] + (h[sp] % np.uint64(5)).astype(np.int64)

    def put(m, seq):
        for k, v in enumerate(seq):
            cols[k][m] = v[m] if isinstance(v, np.ndarray) else v
        for k in range(len(seq), NC):
            cols[k][m] = -1

    m = r < 70  # wiki link
    put(m, [N[b"<a rel=\"mw:WikiLink\" href=\"./"], w, N[b"\" data-parsoid='{\"tsr\":["], p0, p1, p2, N[b","], e0, e1, e2,
            N[b"],\"a\":{\"href\":\"./"], w, N[b"\"},\"sa\":{\"href\":\""], w, N[b"\"},\"stx\":\"piped\",\"dsr\":["], p0, p1, p2, N[b","], e0, e1, e2,
            N[b",2]}'>"], w, N[b"</a>"], N[b" "]])
    m = (r >= 70) & (r < 82)  # bold
    put(m, [N[b"<b data-parsoid='{\"tsr\":["], p0, p1, p2, N[b","], e0, e1, e2, N[b"],\"dsr\":["], p0, p1, p2, N[b","], e0, e1, e2,
            N[b",3,3]}'>"], w, N[b"</b>"], N[b" "]])
    m = (r >= 82) & (r < 92)  # paragraph break
    put(m, [N[b". "], N[b"</p>\n\n<p data-parsoid='{\"dsr\":["], p0, p1, p2, N[b","], e0, e1, e2, N[b",0,0]}'>"], w, N[b" "]])
    m = (r >= 92) & (r < 100)  # citation mark
    cite = T["NUM"] + (pos // 700) % 10000
    put(m, [N[b"<span class=\"reference\" data-parsoid='{\"dsr\":["], p0, p1, p2, N[b","], e0, e1, e2, N[b"]}'"], N[b">"],
            N[b"<a href=\"#cite_note-"], cite, N[b"\">["], cite, N[b"]</a></span>"], N[b" "]])
    ids = np.stack(cols, axis=1).reshape(-1)
    ids = ids[ids >= 0]
    return _assemble(ids, T, CHUNK)


def _gen(n, seed, fn, first_chunk=0, workers=None, out=None):
    """chunks are independent, so they are produced by a few threads (numpy releases the GIL in its array loops); the bytes
    do not depend on the number of workers"""
    import os
    out = np.empty(n, dtype=np.uint8) if out is None else out
    offs = list(range(0, n, CHUNK))
    _table()

    def one(i):
        off = offs[i]
        m = min(CHUNK, n - off)
        out[off:off + m] = fn(seed, first_chunk + i)[:m]

    if workers is None:
        workers = min(8, os.cpu_count() or 1)
    if workers <= 1 or len(offs) < 4:
        for i in range(len(offs)):
            one(i)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(workers) as ex:
            list(ex.map(one, range(len(offs))))
    return out


def gen_text(n, seed=8, first_chunk=0, workers=None, out=None):
    """n bytes of enwik8-like text as a uint8 array; chunk k of the corpus is bytes [k*1e6, (k+1)*1e6)."""
    return _gen(int(n), int(seed), _text_chunk, first_chunk, workers, out)


def gen_html(n, seed=5, first_chunk=0, workers=None, out=None):
    return _gen(int(n), int(seed), _html_chunk, first_chunk, workers, out)


def gen_adversarial(name):
    """SURVEY.md 8d C5b at full size: suffix-sort and cut-walk stress inputs."""
    if name == "zeros1e8":
        return np.zeros(100_000_000, dtype=np.uint8)
    if name == "ab5e7":
        return np.tile(np.frombuffer(b"ab", dtype=np.uint8), 50_000_000)
    if name == "rand1e8":
        with np.errstate(over="ignore"):
            return (_mix(np.arange(12_500_000, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(99))).view(np.uint8)[:100_000_000].copy()
    if name == "line97x1e6":
        line = b"0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ-the quick brown fox jumps over a \n"
        assert len(line) == 97
        return np.tile(np.frombuffer(line, dtype=np.uint8), 1_000_000)
    raise KeyError(name)


ADVERSARIAL = ("zeros1e8", "ab5e7", "rand1e8", "line97x1e6")
