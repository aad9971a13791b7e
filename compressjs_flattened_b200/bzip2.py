"""Host-side mirror of the reference's `Bzip2` object (BJ = /root/reference/Bzip2_joined_.js).

    Bzip2.compressFile(input, output=None, props=None)        BJ:2199-2249
    Bzip2.decompressFile(input, output=None, multistream=False)  BJ:1769-1796
    Bzip2.decompressBlock(input, bitpos, output=None)         BJ:1797-1818
    Bzip2.table(input, callback, multistream=False)           BJ:1823-1863

Same names, argument meaning and error behaviour as the JavaScript API; the work is done by
the CUDA library behind include/bz2b200.h.  Coercions follow Util.coerceInputStream /
coerceOutputStream (BJ:178-272):
  input : bytes-like / list of ints / object with readByte() (-1 at EOF)
  output: None -> returns a new bytes object (the JS returns a fresh Uint8Array);
          int  -> expected size, TypeError('outputsize does not match decoded input') if wrong;
          bytearray/memoryview -> filled in place, same size check;
          object with writeByte(b) -> bytes pushed one at a time, flush() called if present,
          the object itself is returned.
Errors: `Bzip2Error` (a TypeError, like the JS `_throw`, BJ:1384-1391) carrying `.errorCode`;
a bad level raises ValueError('Invalid block size multiplier') (JS: Error, BJ:2208).
"""
import ctypes as C

import numpy as np

from . import _native

EOF = -1


class Bzip2Error(TypeError):
    def __init__(self, code, message):
        super().__init__(message)
        self.errorCode = code


class Err:  # BJ:1365-1375
    OK = 0
    LAST_BLOCK = -1
    NOT_BZIP_DATA = -2
    UNEXPECTED_INPUT_EOF = -3
    UNEXPECTED_OUTPUT_EOF = -4
    DATA_ERROR = -5
    OUT_OF_MEMORY = -6
    OBSOLETE_INPUT = -7
    END_OF_BLOCK = -8


def _coerce_input(inp):
    """Util.coerceInputStream (BJ:178-220): drain any source into a contiguous uint8 array."""
    if hasattr(inp, "readByte"):
        out = bytearray()
        while True:
            ch = inp.readByte()
            if ch == EOF:
                break
            out.append(ch)
        return np.frombuffer(bytes(out), dtype=np.uint8)
    if isinstance(inp, np.ndarray):
        return np.ascontiguousarray(inp, dtype=np.uint8).reshape(-1)
    if isinstance(inp, (bytes, bytearray, memoryview)):
        return np.frombuffer(inp, dtype=np.uint8)
    return np.asarray(list(inp), dtype=np.uint8)


def _deliver(data, output):
    """Util.coerceOutputStream + BufferStream.getBuffer (BJ:222-272)."""
    if output is None or output is False:
        return bytes(data)
    if hasattr(output, "writeByte"):
        for b in data:
            output.writeByte(b)
        if hasattr(output, "flush"):
            output.flush()
        return output
    if isinstance(output, int):
        if output != len(data):
            raise TypeError("outputsize does not match decoded input")
        return bytes(data)
    mv = memoryview(output)
    if len(mv) != len(data):
        raise TypeError("outputsize does not match decoded input")
    mv[:] = data
    return output


class Bzip2Engine:
    """One CUDA context (one GPU).  `Bzip2` below is the process-wide default instance."""

    def __init__(self, device=0, library=None):
        self._lib = library or _native.default_library()
        self._L = self._lib.L
        self._ctx = C.c_void_p()
        rc = self._L.bz2b200_create(device, C.byref(self._ctx))
        if rc:
            raise RuntimeError(f"bz2b200_create(device={device}) failed ({rc}): no usable CUDA device; "
                               "this package has no CPU fallback")
        self.Err = Err

    def close(self):
        if self._ctx:
            self._L.bz2b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- error mapping (BJ:1376-1391) --
    def _raise(self, rc):
        if rc == _native.E_LEVEL:
            raise ValueError("Invalid block size multiplier")
        msg = self._L.bz2b200_strerror(rc).decode()
        if rc in (_native.E_CUDA, _native.E_ARG):
            raise RuntimeError(msg + ": " + self._L.bz2b200_last_error(self._ctx).decode())
        detail = self._L.bz2b200_last_error(self._ctx).decode()
        raise Bzip2Error(rc, msg + (": " + detail if detail else ""))  # _throw(status, optDetail), BJ:1384-1391

    def _take(self, ptr, n):
        data = _native.take_bytes(ptr, n)
        self._L.bz2b200_free(ptr)
        return data

    # -- the four methods of the reference object --
    def compressFile(self, input, output=None, props=None):
        level = props if isinstance(props, (int, float)) and not isinstance(props, bool) else 9  # BJ:2204-2206
        if level < 1 or level > 9 or int(level) != level:
            raise ValueError("Invalid block size multiplier")
        a = _coerce_input(input)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = self._L.bz2b200_compress(self._ctx, a.ctypes.data, a.size, int(level), C.byref(out), C.byref(n))
        if rc:
            self._raise(rc)
        return _deliver(self._take(out, n.value), output)

    # -- the stream flavour (what NPM/bin/compressjs:163-180 does with {readByte}/{writeByte} streams) --
    @staticmethod
    def _reader(source, piece):
        if hasattr(source, "read"):
            return source.read
        if hasattr(source, "readByte"):
            def read(n):
                out = bytearray()
                while len(out) < n:
                    ch = source.readByte()
                    if ch == EOF:
                        break
                    out.append(ch)
                return bytes(out)
            return read
        view, pos = memoryview(_coerce_input(source)), [0]

        def read(n):
            b = view[pos[0]:pos[0] + n].tobytes()
            pos[0] += len(b)
            return b
        return read

    def _pump(self, kind, handle, source, sink, piece):
        """read pieces, feed them, hand on what comes back; the C stream object does the work (csrc/stream_abi.inl)"""
        L = self._L
        feed, finish = getattr(L, f"bz2b200_{kind}_feed"), getattr(L, f"bz2b200_{kind}_finish")
        read = self._reader(source, piece)
        collected = bytearray() if sink is None else None

        def put(ptr, n):
            data = self._take(ptr, n) if n else b""
            if not data:
                return
            if collected is not None:
                collected.extend(data)
            elif hasattr(sink, "write"):
                sink.write(data)
            else:
                for x in data:
                    sink.writeByte(x)

        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        try:
            while True:
                data = read(piece)
                if not data:
                    break
                a = np.frombuffer(data, dtype=np.uint8)
                rc = feed(handle, a.ctypes.data, a.size, C.byref(out), C.byref(n))
                if rc:
                    self._raise(rc)
                put(out, n.value)
                if len(data) < piece:
                    break
            rc = finish(handle, C.byref(out), C.byref(n))
            if rc:
                self._raise(rc)
            put(out, n.value)
        finally:
            getattr(L, f"bz2b200_{kind}_close")(handle)
        if sink is None:
            return bytes(collected)
        if hasattr(sink, "flush"):
            sink.flush()
        return sink

    def compressStream(self, source, sink=None, props=None, chunk_bytes=64 << 20, piece_bytes=None):
        """Stream flavour of compressFile: the source is read piece by piece and whole blocks are compressed as soon as the
        bytes after them (their halo) have arrived, so neither the input nor the output is ever held in full.
        Byte-identical to compressFile on the concatenated input.  source: .read(n) -> bytes, {readByte}, or bytes-like;
        sink: .write(bytes), {writeByte}, or None (the stream is returned as bytes)."""
        level = props if isinstance(props, (int, float)) and not isinstance(props, bool) else 9  # BJ:2204-2206
        if level < 1 or level > 9 or int(level) != level:
            raise ValueError("Invalid block size multiplier")                                    # BJ:2208-2210
        h = C.c_void_p()
        rc = self._L.bz2b200_zstream_open(self._ctx, int(level), int(chunk_bytes), C.byref(h))
        if rc:
            self._raise(rc)
        return self._pump("zstream", h, source, sink, int(piece_bytes or max(1, chunk_bytes // 4)))

    def decompressStream(self, source, sink=None, multistream=False, chunk_bytes=16 << 20, piece_bytes=None):
        """Stream flavour of decompressFile: blocks are decoded as soon as they are complete; the bytes of the blocks before
        an error have been delivered when it is raised (like the reference, which writes as it goes)."""
        h = C.c_void_p()
        rc = self._L.bz2b200_dstream_open(self._ctx, int(bool(multistream)), int(chunk_bytes), C.byref(h))
        if rc:
            self._raise(rc)
        return self._pump("dstream", h, source, sink, int(piece_bytes or max(1, chunk_bytes // 4)))

    def decompressFile(self, input, output=None, multistream=False):
        a = _coerce_input(input)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = self._L.bz2b200_decompress(self._ctx, a.ctypes.data, a.size, int(bool(multistream)), C.byref(out), C.byref(n))
        if rc:
            self._raise(rc)
        return _deliver(self._take(out, n.value), output)

    def decompressBlock(self, input, pos, output=None):
        a = _coerce_input(input)
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = self._L.bz2b200_decompress_block(self._ctx, a.ctypes.data, a.size, int(pos), C.byref(out), C.byref(n))
        if rc:
            self._raise(rc)
        return _deliver(self._take(out, n.value), output)

    def table(self, input, callback, multistream=False):
        a = _coerce_input(input)
        pos, sz, n = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.c_size_t()
        rc = self._L.bz2b200_table(self._ctx, a.ctypes.data, a.size, int(bool(multistream)), C.byref(pos), C.byref(sz), C.byref(n))
        if rc:
            self._raise(rc)
        try:
            for i in range(n.value):
                callback(int(pos[i]), int(sz[i]))
        finally:
            self._L.bz2b200_free(pos)
            self._L.bz2b200_free(sz)

    # -- extras used by bench.py / tests --
    def stats(self):
        st = _native.Stats()
        self._L.bz2b200_last_stats(self._ctx, C.byref(st))
        return st

    def compress_device(self, d_in_ptr, n, level, d_out_ptr, out_cap):
        olen = C.c_size_t()
        rc = self._L.bz2b200_compress_device(self._ctx, d_in_ptr, n, level, d_out_ptr, out_cap, C.byref(olen))
        if rc:
            self._raise(rc)
        return olen.value

    def decompress_device(self, d_in_ptr, n, multistream, d_out_ptr, out_cap):
        olen = C.c_size_t()
        rc = self._L.bz2b200_decompress_device(self._ctx, d_in_ptr, n, int(bool(multistream)), d_out_ptr, out_cap, C.byref(olen))
        if rc:
            self._raise(rc)
        return olen.value

    def compress_bound(self, n, level=9):
        return int(self._L.bz2b200_compress_bound(n, level))

    def debug_fetch(self, what, blk, nbytes):
        buf = np.zeros(nbytes, dtype=np.uint8)
        got = self._L.bz2b200_debug_fetch(self._ctx, what, blk, buf.ctypes.data, nbytes)
        if got < 0:
            self._raise(int(got))
        return buf[:got]

    # -- block-range shards (include/bz2b200.h, "multi-GPU") --
    def shard_begin(self, data, level, device_ptr=None, nbytes=None):
        if device_ptr is not None:
            rc = self._L.bz2b200_shard_begin(self._ctx, device_ptr, nbytes, 1, level)
        else:
            a = _coerce_input(data)
            self._shard_keep = a
            rc = self._L.bz2b200_shard_begin(self._ctx, a.ctypes.data, a.size, 0, level)
        if rc:
            self._raise(rc)

    def shard_cut(self, s_start, own_len, is_last):
        info = _native.ShardInfo()
        rc = self._L.bz2b200_shard_cut(self._ctx, s_start, own_len, int(is_last), C.byref(info))
        if rc:
            self._raise(rc)
        return info

    def shard_gtotal(self, pos):
        g = C.c_uint64()
        rc = self._L.bz2b200_shard_gtotal(self._ctx, pos, C.byref(g))
        if rc:
            self._raise(rc)
        return g.value

    def shard_cut_g(self, g_before, own_len):
        """Speculative cut walks (g_before = G of the earlier shards); shard_cut_pick chooses among them."""
        rc = self._L.bz2b200_shard_cut_g(self._ctx, g_before, own_len)
        if rc:
            self._raise(rc)

    def shard_cut_pick(self, s_start, own_len, is_last):
        """The speculated walk that started at s_start, or None (then call shard_cut)."""
        info, found = _native.ShardInfo(), C.c_int()
        rc = self._L.bz2b200_shard_cut_pick(self._ctx, s_start, own_len, int(is_last), C.byref(info), C.byref(found))
        if rc:
            self._raise(rc)
        return info if found.value else None

    def shard_compress(self, info):
        rc = self._L.bz2b200_shard_compress(self._ctx, C.byref(info))
        if rc:
            self._raise(rc)
        return info

    def shard_emit(self, info, bit_phase, to_host=True):
        """to_host: True -> bytes; False -> the segment stays in HBM, its length is returned;
        "raw" -> (pointer, length) in the library's page-locked result memory, released with free_raw."""
        out, n = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = self._L.bz2b200_shard_emit(self._ctx, bit_phase, C.byref(info), C.byref(out) if to_host else None, C.byref(n))
        if rc:
            self._raise(rc)
        if to_host == "raw":
            return out, n.value
        return self._take(out, n.value) if to_host else n.value

    def free_raw(self, ptr):
        self._L.bz2b200_free(ptr)

    def stitch_shards(self, level, segs, infos):
        n = len(segs)
        arr = (C.c_char_p * n)(*[bytes(s) if len(s) else b"\0" for s in segs])
        inf = (_native.ShardInfo * n)(*infos)
        out, ln = C.POINTER(C.c_uint8)(), C.c_size_t()
        rc = self._L.bz2b200_stitch_shards(level, n, arr, inf, C.byref(out), C.byref(ln))
        if rc:
            self._raise(rc)
        return self._take(out, ln.value)

    def debug_set_block_cap(self, cap):
        """tests only: 0 restores level*100000-19"""
        rc = self._L.bz2b200_debug_set_block_cap(self._ctx, cap)
        if rc:
            self._raise(rc)

    def debug_set_batch_blocks(self, blocks):
        """tests only: blocks per batch of the per-block stages, 0 restores the automatic limit"""
        rc = self._L.bz2b200_debug_set_batch_blocks(self._ctx, blocks)
        if rc:
            self._raise(rc)

    def debug_set_ignore_block_crc(self, on):
        """tests only: decode without comparing block CRCs (what a damaged stream decodes TO)"""
        rc = self._L.bz2b200_debug_set_ignore_block_crc(self._ctx, int(bool(on)))
        if rc:
            self._raise(rc)

    def debug_huffman_lengths(self, sorted_freqs, maxlen):
        """tests only: HuffmanAllocator.allocateHuffmanCodeLengths (BJ:1275-1298) run by the device copy of the allocator"""
        a = np.asarray(sorted_freqs, dtype=np.int32).copy()
        rc = self._L.bz2b200_debug_huffman_lengths(self._ctx, a.ctypes.data, a.size, int(maxlen))
        if rc:
            self._raise(rc)
        return [int(x) for x in a]

    def debug_set_pool(self, min_bytes=32_000_000, shard_bytes=0, first_halo=0, force_staging=False):
        """tests only: which inputs compressFile sends through the context's two-lane pool, and how they are cut"""
        rc = self._L.bz2b200_debug_set_pool(self._ctx, min_bytes, shard_bytes, first_halo, int(bool(force_staging)))
        if rc:
            self._raise(rc)

    def block_table(self):
        nb = self.stats().n_blocks
        raw = self.debug_fetch(0, 0, C.sizeof(_native.BlockRec) * nb)
        return (_native.BlockRec * nb).from_buffer_copy(raw.tobytes()) if nb else []

    def block_meta(self):
        nb = self.stats().n_blocks
        raw = self.debug_fetch(4, 0, C.sizeof(_native.BlockMeta) * nb)
        return (_native.BlockMeta * nb).from_buffer_copy(raw.tobytes()) if nb else []


class _Lazy:
    """`Bzip2` global of the reference (BJ:2198-2255): created on first use."""
    _eng = None

    def _get(self):
        if _Lazy._eng is None:
            _Lazy._eng = Bzip2Engine(0)
        return _Lazy._eng

    def __getattr__(self, name):
        return getattr(self._get(), name)


Bzip2 = _Lazy()
