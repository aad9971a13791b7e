// bzip2_shim.js -- drop-in for the globals that Bzip2_joined_.js defines (BJ:3-10): `Bzip2` does its work in the CUDA
// library through the N-API addon; `Stream`, `BitStream`, `Util`, `CRC32` are small stand-ins written for this file
// (enough for callers that build {readByte}/{writeByte} streams the way NPM/bin/compressjs:60-135 does); `BWT` and
// `HuffmanAllocator` are the internals the GPU path replaces -- if the original script was loaded first its objects
// are kept, otherwise touching them says so.
//
// Load it exactly like the joined script: vm.runInThisContext(fs.readFileSync(...)), a <script> tag in an
// Electron/NW.js context, or require().  A script run by vm.runInThisContext has NO `require` in scope, so nothing
// here calls a bare require: the addon is taken from globalThis.BZ2B200_ADDON if the host set it, else loaded with
// whatever loader the context offers (module.require, process.mainModule.require, a global require), on first use.
//
// One algorithm of the reference IS restated in the native library (not in this file): the in-place length-limited
// code-length allocator (csrc/huff.cuh ha_first / ha_allocate follow BJ:1135-1298 statement by statement, because the
// output must be byte-identical).  Everything else is an independent design.
"use strict";
var Bzip2 = (function (root) {
  var EOF = -1, addon = null;
  var Err = { OK: 0, LAST_BLOCK: -1, NOT_BZIP_DATA: -2, UNEXPECTED_INPUT_EOF: -3, UNEXPECTED_OUTPUT_EOF: -4,
              DATA_ERROR: -5, OUT_OF_MEMORY: -6, OBSOLETE_INPUT: -7, END_OF_BLOCK: -8 };
  var PIECE = 16 << 20;                    // bytes drained from a {readByte} source per feed

  function native() {
    if (addon) return addon;
    if (root.BZ2B200_ADDON) return (addon = root.BZ2B200_ADDON);
    var path = root.BZ2B200_ADDON_PATH || './build/Release/bz2b200_napi.node';
    var load = (typeof module !== 'undefined' && module && typeof module.require === 'function') ? module.require.bind(module)
             : (typeof process !== 'undefined' && process.mainModule && process.mainModule.require) ? process.mainModule.require.bind(process.mainModule)
             : (typeof require === 'function') ? require : null;
    if (!load) throw new Error('bz2b200: no module loader in this context; set globalThis.BZ2B200_ADDON = require(".../bz2b200_napi.node") before the first call');
    return (addon = load(path));
  }
  function isSource(x) { return x && typeof x === 'object' && typeof x.readByte === 'function'; }
  function isSink(x) { return x && typeof x === 'object' && typeof x.writeByte === 'function'; }
  function toBytes(input) {                // Util.coerceInputStream (BJ:178-220), for indexable inputs
    if (input instanceof Uint8Array) return input;   // Buffer is a Uint8Array
    return Uint8Array.from(input);                    // plain Array
  }
  function readPiece(src, n) {             // up to n bytes from a {readByte} source (uses read() when the source has one)
    var buf = new Uint8Array(n), got = 0, ch;
    if (typeof src.read === 'function') {
      while (got < n) { var r = src.read(buf, got, n - got); if (!(r > 0)) break; got += r; }
    } else {
      while (got < n && (ch = src.readByte()) !== EOF) buf[got++] = ch;
    }
    return got === n ? buf : buf.subarray(0, got);
  }
  function drain(src) {                    // a whole {readByte} source (decompressBlock / table need random access)
    var parts = [], total = 0, piece;
    do { piece = readPiece(src, 1 << 20); parts.push(piece); total += piece.length; } while (piece.length === (1 << 20));
    var out = new Uint8Array(total), o = 0;
    parts.forEach(function (p) { out.set(p, o); o += p.length; });
    return out;
  }
  function check(r) {                      // _throw (BJ:1384-1391)
    if (r.rc === 0) return r;
    if (r.rc === -100) throw new Error('Invalid block size multiplier');
    var e = (r.rc <= -101) ? new Error(r.message) : new TypeError(r.message);
    e.errorCode = r.rc;
    throw e;
  }
  function Collector(output) {             // Util.coerceOutputStream + BufferStream.getBuffer (BJ:222-272)
    this.sink = isSink(output) ? output : null;
    this.output = output;
    this.parts = [];
    this.total = 0;
  }
  Collector.prototype.put = function (data) {
    if (!data || !data.length) return;
    this.total += data.length;
    if (this.sink) { for (var i = 0; i < data.length; i++) this.sink.writeByte(data[i]); }
    else this.parts.push(data);
  };
  Collector.prototype.result = function () {
    if (this.sink) { if (this.sink.flush) this.sink.flush(); return this.sink; }
    var out = this.output;
    var fixed = (typeof out === 'number') ? out : (out && typeof out === 'object' && 'length' in out) ? out.length : -1;
    if (fixed >= 0 && fixed !== this.total) throw new TypeError('outputsize does not match decoded input');
    var dst = (out && typeof out === 'object' && 'length' in out) ? out
            : (this.parts.length === 1) ? this.parts[0] : new Uint8Array(this.total);
    if (dst !== this.parts[0]) { var o = 0; this.parts.forEach(function (p) { dst.set ? dst.set(p, o) : p.forEach(function (b, k) { dst[o + k] = b; }); o += p.length; }); }
    return dst;
  };
  // a {readByte} source is pumped through the addon's stream objects (bz2b200_zstream_* / bz2b200_dstream_*), so neither
  // the input nor the output is held in full (the reference's stream flavour, NPM/bin/compressjs:163-180)
  function pump(handle, src, col) {
    var a = native();
    try {
      for (;;) {
        var piece = readPiece(src, PIECE);
        if (piece.length) col.put(check(a.streamFeed(handle, piece)).data);
        if (piece.length < PIECE) break;
      }
      col.put(check(a.streamFinish(handle)).data);
    } finally { a.streamClose(handle); }
    return col.result();
  }

  var B = Object.create(null);
  B.Err = Err;
  B.compressFile = function (inStream, outStream, props) {
    var level = (typeof props === 'number') ? props : 9;               // BJ:2204-2206
    if (level < 1 || level > 9) throw new Error('Invalid block size multiplier');
    var col = new Collector(outStream), a = native();
    if (isSource(inStream)) return pump(check(a.zstreamOpen(level)).handle, inStream, col);
    col.put(check(a.compress(toBytes(inStream), level)).data);
    return col.result();
  };
  B.decompressFile = function (input, output, multistream) {
    var col = new Collector(output), a = native();
    if (isSource(input)) return pump(check(a.dstreamOpen(!!multistream)).handle, input, col);
    var hint = (typeof output === 'number') ? output : 0;              // the expected size doubles as the allocation hint
    col.put(check(a.decompress(toBytes(input), !!multistream, hint)).data);
    return col.result();
  };
  B.decompressBlock = function (input, pos, output) {
    var col = new Collector(output), bytes = isSource(input) ? drain(input) : toBytes(input);
    col.put(check(native().decompressBlock(bytes, pos)).data);
    return col.result();
  };
  B.table = function (input, callback, multistream) {
    var bytes = isSource(input) ? drain(input) : toBytes(input);
    check(native().table(bytes, !!multistream)).table.forEach(function (row) { callback(row[0], row[1]); });
  };
  // every GPU of the box behind the same call (bz2b200_pool_*): Bzip2.useDevices([0,1,2,3]) then compressFile as before
  B.useDevices = function (devices, lanesPerDevice) { check(native().useDevices(devices, lanesPerDevice || 1)); };
  return B;
}(typeof globalThis !== 'undefined' ? globalThis : this));

// ---- the sibling globals of the joined script (BJ:3-10) ----
var Stream = (typeof Stream !== 'undefined') ? Stream : (function () {      // the abstract base of BJ:12-62
  function S() {}
  S.prototype.readByte = function () { throw new Error('abstract method readByte() not implemented'); };
  S.prototype.read = function (buf, off, len) {
    var n = 0, ch;
    while (n < len && (ch = this.readByte()) !== -1) buf[off + n++] = ch;
    return n;
  };
  S.prototype.writeByte = function () { throw new Error('abstract method writeByte() not implemented'); };
  S.prototype.write = function (buf, off, len) { for (var i = 0; i < len; i++) this.writeByte(buf[off + i]); return len; };
  S.prototype.flush = function () {};
  S.prototype.seek = function () { throw new Error('abstract method seek() not implemented'); };
  S.prototype.tell = function () { throw new Error('abstract method tell() not implemented'); };
  S.prototype.eof = function () { throw new Error('abstract method eof() not implemented'); };
  S.EOF = -1;
  return S;
}());
var CRC32 = (typeof CRC32 !== 'undefined') ? CRC32 : (function () {         // bzip2's MSB-first CRC-32 (BJ:1013-1079)
  var T = new Uint32Array(256);
  for (var i = 0; i < 256; i++) { var c = i << 24; for (var k = 0; k < 8; k++) c = (c & 0x80000000) ? ((c << 1) ^ 0x04C11DB7) : (c << 1); T[i] = c >>> 0; }
  function C() { var crc = 0xFFFFFFFF;
    this.getCRC = function () { return (~crc) >>> 0; };
    this.updateCRC = function (v) { crc = ((crc << 8) ^ T[((crc >>> 24) ^ v) & 0xFF]) >>> 0; };
    this.updateCRCRun = function (v, n) { while (n-- > 0) this.updateCRC(v); };
  }
  return C;
}());
var Util = (typeof Util !== 'undefined') ? Util : {
  EOF: -1,
  makeU8Buffer: function (n) { return new Uint8Array(n); },
  makeU16Buffer: function (n) { return new Uint16Array(n); },
  makeU32Buffer: function (n) { return new Uint32Array(n); },
  makeS32Buffer: function (n) { return new Int32Array(n); },
  fls: function (v) { var r = 0; v = v >>> 0; while (v) { r++; v >>>= 1; } return r; },          // BJ:470-486
  log2c: function (v) { return v === 0 ? -1 : Util.fls(v - 1); }
};
var BitStream = (typeof BitStream !== 'undefined') ? BitStream : (function () {   // MSB-first bit reader/writer (BJ:64-166)
  function BS(stream) { this.stream = stream; this.bitOffset = 0; this.curByte = 0; this.wbits = 0; this.wbyte = 0; }
  BS.prototype.readBits = function (n) {
    var v = 0;
    while (n-- > 0) {
      if (this.bitOffset === 0) { var b = this.stream.readByte(); this.curByte = b === -1 ? 0 : b; this.bitOffset = 8; }
      v = v * 2 + ((this.curByte >>> --this.bitOffset) & 1);
    }
    return v;
  };
  BS.prototype.writeBits = function (n, v) {
    for (var i = n - 1; i >= 0; i--) {
      this.wbyte = (this.wbyte << 1) | (Math.floor(v / Math.pow(2, i)) & 1);
      if (++this.wbits === 8) { this.stream.writeByte(this.wbyte); this.wbits = 0; this.wbyte = 0; }
    }
  };
  BS.prototype.flush = function () { while (this.wbits) this.writeBits(1, 0); if (this.stream.flush) this.stream.flush(); };
  return BS;
}());
var BWT = (typeof BWT !== 'undefined') ? BWT : new Proxy({}, { get: function (_, k) {
  throw new Error('BWT.' + String(k) + ': the suffix sort lives on the GPU in this build (csrc/bwt.cuh); load the original Bzip2_joined_.js first if you need the JavaScript one'); } });
var HuffmanAllocator = (typeof HuffmanAllocator !== 'undefined') ? HuffmanAllocator : new Proxy({}, { get: function (_, k) {
  throw new Error('HuffmanAllocator.' + String(k) + ': the code-length allocator lives on the GPU in this build (csrc/huff.cuh); load the original Bzip2_joined_.js first if you need the JavaScript one'); } });
if (typeof module === 'object' && module && module.exports) module.exports = Bzip2;   // require() users get the object too
