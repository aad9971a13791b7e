"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ORACLE_DIR, "liboracle.so")

SORT_STABLE, SORT_LEGACY_V8 = 0, 1


class Stats(C.Structure):
    _fields_ = [("n_blocks", C.c_uint32), ("d1_triggered", C.c_uint32),
                ("in_bytes", C.c_uint64), ("out_bytes", C.c_uint64),
                ("rle1_bytes", C.c_uint64), ("mtf_syms", C.c_uint64)]


class BlockInfo(C.Structure):
    _fields_ = [("n", C.c_uint32), ("orig_ptr", C.c_uint32), ("alpha", C.c_uint32),
                ("m", C.c_uint32), ("n_groups", C.c_uint32), ("n_sel", C.c_uint32),
                ("d1", C.c_uint32), ("bits", C.c_uint64)]


def build(force=False):
    src = [os.path.join(_ORACLE_DIR, f) for f in ("bz2_oracle.c", "bz2_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, szp = C.POINTER(C.c_uint8), C.POINTER(C.c_size_t)
        L.orc_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(u8p), szp, C.POINTER(Stats)]
        L.orc_compress_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.POINTER(u8p), szp, C.POINTER(Stats)]
        L.orc_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(u8p), szp]
        L.orc_decompress_mt.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(u8p), szp]
        L.orc_decompress_block.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(u8p), szp]
        L.orc_table.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.POINTER(C.c_uint64)),
                                C.POINTER(C.POINTER(C.c_uint32)), szp]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_strerror.restype = C.c_char_p
        L.orc_crc32.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_crc32.restype = C.c_uint32
        L.orc_fls.argtypes = [C.c_uint64]
        L.orc_rle1_block.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, szp, C.POINTER(C.c_uint32)]
        L.orc_rle1_block.restype = C.c_size_t
        L.orc_cut_points.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.POINTER(C.c_uint64)),
                                     C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_uint32))]
        L.orc_cut_points.restype = C.c_size_t
        L.orc_bwt.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_huff_alloc.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_huff_alloc.restype = None
        L.orc_huff_lengths.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_huff_lengths.restype = None
        L.orc_block_stages.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(BlockInfo),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_debug_set_block_cap.argtypes = [C.c_size_t]
        L.orc_debug_set_block_cap.restype = None
        L.orc_debug_set_ignore_block_crc.argtypes = [C.c_int]
        L.orc_debug_set_ignore_block_crc.restype = None
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, rc):
        self.errorCode = rc
        super().__init__(lib().orc_strerror(rc).decode())


def _as_u8(buf):
    a = np.frombuffer(bytes(buf), dtype=np.uint8) if not isinstance(buf, np.ndarray) else np.ascontiguousarray(buf, dtype=np.uint8)
    return a


def _take(ptr, n):
    if n >= 1 << 31:   # ctypes.string_at takes an int-sized length
        out = bytes(memoryview((C.c_ubyte * n).from_address(C.addressof(ptr.contents))))
    else:
        out = C.string_at(ptr, n) if n else b""
    lib().orc_free(ptr)
    return out


def compress(data, level=9, sort_mode=SORT_STABLE, threads=1, return_stats=False):
    a = _as_u8(data)
    out, n, st = C.POINTER(C.c_uint8)(), C.c_size_t(), Stats()
    rc = lib().orc_compress_mt(a.ctypes.data, a.size, level, sort_mode, threads, C.byref(out), C.byref(n), C.byref(st))
    if rc:
        raise OracleError(rc)
    res = _take(out, n.value)
    return (res, st) if return_stats else res


def decompress(data, multistream=False, threads=1):
    a = _as_u8(data)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = lib().orc_decompress_mt(a.ctypes.data, a.size, int(multistream), threads, C.byref(out), C.byref(n))
    if rc:
        raise OracleError(rc)
    return _take(out, n.value)


def decompress_block(data, bitpos):
    a = _as_u8(data)
    out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
    rc = lib().orc_decompress_block(a.ctypes.data, a.size, bitpos, C.byref(out), C.byref(n))
    if rc:
        raise OracleError(rc)
    return _take(out, n.value)


def table(data, multistream=False):
    a = _as_u8(data)
    pos, sz, n = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.c_size_t()
    rc = lib().orc_table(a.ctypes.data, a.size, int(multistream), C.byref(pos), C.byref(sz), C.byref(n))
    if rc:
        raise OracleError(rc)
    res = [(int(pos[i]), int(sz[i])) for i in range(n.value)]
    lib().orc_free(pos)
    lib().orc_free(sz)
    return res


def set_block_cap(cap):
    """test hook: 0 restores the reference's capacity"""
    lib().orc_debug_set_block_cap(cap)


def set_ignore_block_crc(on):
    """test hook: decode damaged streams without the block CRC comparison"""
    lib().orc_debug_set_ignore_block_crc(1 if on else 0)


def crc32(data):
    a = _as_u8(data)
    return int(lib().orc_crc32(a.ctypes.data, a.size))


def fls(v):
    return int(lib().orc_fls(v))


def rle1_block(data, cap):
    a = _as_u8(data)
    blk = np.zeros(cap, dtype=np.uint8)
    used, crc = C.c_size_t(), C.c_uint32()
    n = lib().orc_rle1_block(a.ctypes.data, a.size, cap, blk.ctypes.data, C.byref(used), C.byref(crc))
    return blk[:n].copy(), int(used.value), int(crc.value)


def cut_points(data, level):
    a = _as_u8(data)
    s, l, c = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)()
    nb = lib().orc_cut_points(a.ctypes.data, a.size, level, C.byref(s), C.byref(l), C.byref(c))
    starts = [int(s[i]) for i in range(nb + 1)]
    lens = [int(l[i]) for i in range(nb)]
    crcs = [int(c[i]) for i in range(nb)]
    for p in (s, l, c):
        lib().orc_free(p)
    return starts, lens, crcs


def bwt(data):
    a = _as_u8(data)
    out = np.zeros(max(a.size, 1), dtype=np.uint8)
    pidx = lib().orc_bwt(a.ctypes.data, a.size, out.ctypes.data)
    return out[:a.size].tobytes(), int(pidx)


def huff_alloc(sorted_freqs, maxlen):
    a = np.array(sorted_freqs, dtype=np.int32)
    lib().orc_huff_alloc(a.ctypes.data, a.size, maxlen)
    return a.tolist()


def huff_lengths(freq):
    f = np.array(freq, dtype=np.int32)
    out = np.zeros(f.size, dtype=np.uint8)
    lib().orc_huff_lengths(f.ctypes.data, f.size, out.ctypes.data)
    return out


def block_stages(block, sort_mode=SORT_STABLE):
    """compressBlock's intermediates for one RLE1'd block."""
    a = _as_u8(block)
    n = a.size
    info = BlockInfo()
    U = np.zeros(n, dtype=np.uint8)
    A = np.zeros(n + 1, dtype=np.uint16)
    sel = np.zeros((n + 1 + 49) // 50 + 1, dtype=np.uint8)
    lens = np.zeros((6, 258), dtype=np.uint8)
    lib().orc_block_stages(a.ctypes.data, n, sort_mode, C.byref(info), U.ctypes.data, A.ctypes.data,
                           sel.ctypes.data, lens.ctypes.data)
    return dict(n=info.n, orig_ptr=info.orig_ptr, alpha=info.alpha, m=info.m, n_groups=info.n_groups,
                n_sel=info.n_sel, d1=info.d1, bits=info.bits, U=U, A=A[:info.m].copy(),
                sel=sel[:info.n_sel].copy(), lens=lens[:info.n_groups, :info.alpha + 2].copy())
