// bzip2_shim.js -- drop-in for the `Bzip2` global that Bzip2_joined_.js defines (written from scratch; the
// reference is GPL/LGPL and none of its code is reused).  Load it exactly like the joined script
// (vm.runInThisContext / <script> in an Electron-style Node context); it leaves a global `Bzip2` with
// compressFile / decompressFile / decompressBlock / table that call the CUDA library through the addon.
// Argument coercions and error shapes follow BJ:178-272, BJ:1365-1391, BJ:2199-2210.
"use strict";
var Bzip2 = (function () {
  var addon = require('./build/Release/bz2b200_napi.node');
  var EOF = -1;
  var Err = { OK: 0, LAST_BLOCK: -1, NOT_BZIP_DATA: -2, UNEXPECTED_INPUT_EOF: -3, UNEXPECTED_OUTPUT_EOF: -4,
              DATA_ERROR: -5, OUT_OF_MEMORY: -6, OBSOLETE_INPUT: -7, END_OF_BLOCK: -8 };

  function toBytes(input) {               // Util.coerceInputStream (BJ:178-220)
    if (input && typeof input === 'object' && 'readByte' in input) {
      var chunks = [], buf = new Uint8Array(65536), n = 0, ch;
      while ((ch = input.readByte()) !== EOF) {
        if (n === buf.length) { chunks.push(buf); buf = new Uint8Array(65536); n = 0; }
        buf[n++] = ch;
      }
      chunks.push(buf.subarray(0, n));
      var total = chunks.reduce(function (a, c) { return a + c.length; }, 0), out = new Uint8Array(total), o = 0;
      chunks.forEach(function (c) { out.set(c, o); o += c.length; });
      return out;
    }
    if (input instanceof Uint8Array) return input;   // Buffer is a Uint8Array
    return Uint8Array.from(input);                    // plain Array
  }
  function deliver(data, output) {         // Util.coerceOutputStream + BufferStream.getBuffer (BJ:222-272)
    if (!output) return data;
    if (typeof output === 'object' && 'writeByte' in output) {
      for (var i = 0; i < data.length; i++) output.writeByte(data[i]);
      if (output.flush) output.flush();
      return output;
    }
    var size = (typeof output === 'number') ? output : output.length;
    if (size !== data.length) throw new TypeError('outputsize does not match decoded input');
    if (typeof output === 'number') return data;
    for (var j = 0; j < data.length; j++) output[j] = data[j];
    return output;
  }
  function check(r) {                      // _throw (BJ:1384-1391)
    if (r.rc === 0) return r;
    if (r.rc === -100) throw new Error('Invalid block size multiplier');
    var e = (r.rc <= -101) ? new Error(r.message) : new TypeError(r.message);
    e.errorCode = r.rc;
    throw e;
  }
  var B = Object.create(null);
  B.Err = Err;
  B.compressFile = function (inStream, outStream, props) {
    var level = (typeof props === 'number') ? props : 9;               // BJ:2204-2206
    if (level < 1 || level > 9) throw new Error('Invalid block size multiplier');
    return deliver(check(addon.compress(toBytes(inStream), level)).data, outStream);
  };
  B.decompressFile = function (input, output, multistream) {
    return deliver(check(addon.decompress(toBytes(input), !!multistream)).data, output);
  };
  B.decompressBlock = function (input, pos, output) {
    return deliver(check(addon.decompressBlock(toBytes(input), pos)).data, output);
  };
  B.table = function (input, callback, multistream) {
    check(addon.table(toBytes(input), !!multistream)).table.forEach(function (row) { callback(row[0], row[1]); });
  };
  return B;
}());
