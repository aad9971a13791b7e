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

    def compressStream(self, source, sink=None, props=None, chunk_bytes=64 << 20):
        """Stream flavour of compressFile (what NPM/bin/compressjs:163-180 does with {readByte}/{writeByte}
        streams): the source is read chunk by chunk and whole blocks are compressed as soon as the bytes after them (their
        halo) have arrived, so neither the input nor the output is ever held in full.  Byte-identical to compressFile on
        the concatenated input.  source: .read(n) -> bytes, or {readByte}; sink: .write(bytes), {writeByte}, or None
        (the stream is returned as bytes).  Built on the shard calls: every chunk is a shard of the stream."""
        level = props if isinstance(props, (int, float)) and not isinstance(props, bool) else 9  # BJ:2204-2206
        if level < 1 or level > 9 or int(level) != level:
            raise ValueError("Invalid block size multiplier")                                    # BJ:2208-2210
        level = int(level)
        if hasattr(source, "read"):
            read = source.read
        elif hasattr(source, "readByte"):
            def read(n):
                out = bytearray()
                while len(out) < n:
                    ch = source.readByte()
                    if ch == EOF:
                        break
                    out.append(ch)
                return bytes(out)
        else:
            view, pos = memoryview(_coerce_input(source)), [0]

            def read(n):
                b = view[pos[0]:pos[0] + n].tobytes()
                pos[0] += len(b)
                return b
        collected = bytearray() if sink is None else None

        def put(b):
            if not b:
                return
            if collected is not None:
                collected.extend(b)
            elif hasattr(sink, "write"):
                sink.write(bytes(b))
            else:
                for x in b:
                    sink.writeByte(x)

        put(b"BZh" + bytes([0x30 + level]))                                                     # BJ:2223-2226
        bitpos, tail, crc, buf, eof = 32, 0, 0, bytearray(), False
        while True:
            if not eof:
                data = read(chunk_bytes)
                eof = len(data) < chunk_bytes
                buf += data
            own = len(buf) if eof else len(buf) // 2          # the second half is the halo of the blocks of the first
            if own or eof:
                self.shard_begin(bytes(buf), level)
                info = self.shard_cut(0, own, eof)
                if not info.complete:                         # a block needs input that is not here yet: read on
                    if eof:
                        raise RuntimeError("internal: incomplete block at end of input")
                    continue
                if info.n_blocks:
                    self.shard_compress(info)
                    seg = self.shard_emit(info, bitpos & 7)
                    merged = bytearray(seg)
                    merged[0] |= tail
                    end = bitpos + int(info.bits)
                    nfull = (end >> 3) - (bitpos >> 3)        # bytes that are complete now
                    put(merged[:nfull])
                    tail = merged[nfull] if end & 7 else 0
                    bitpos = end
                    m = info.n_blocks & 31
                    crc = (((crc << m) | (crc >> (32 - m))) & 0xFFFFFFFF if m else crc) ^ info.crc_fold   # BJ:2237
                    del buf[:int(info.next_start)]
            if eof:
                break
        foot = (0x177245385090 << 32) | crc                                                     # BJ:2245-2247
        nbits = (bitpos & 7) + 80                     # the pending bits of the last byte, then the footer
        pad = -nbits % 8                              # zero bits up to the byte boundary (BitStream.flush, BJ:127-132)
        val = (tail << (nbits + pad - 8)) | (foot << pad) if bitpos & 7 else foot
        put(val.to_bytes((nbits + pad) // 8, "big"))
        if sink is None:
            return bytes(collected)
        if hasattr(sink, "flush"):
            sink.flush()
        return sink

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
