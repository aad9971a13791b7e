"""Command line of the bzip2 path, with the flags and defaults of the reference's bin/compressjs
(NPM/bin/compressjs:7-33,58,163-180):

  python -m compressjs_flattened_b200 -d|-z [-1 .. -9] [-b <bits>] [-t bzip2] [infile] [outfile]

  -z  compress (the default), stream flavour: the input is read piece by piece (bz2b200_zstream_*), level 7 by default
  -d  decompress the FIRST stream of the input, stream flavour (bz2b200_dstream_*); --multistream decodes all of them
      (an extension: the reference's command line never passes the flag, NPM/bin/compressjs:166)
  -b  with -d: extract the single block whose signature starts at bit <bits> (Bzip2.decompressBlock)
If <infile> is omitted, reads from stdin; if <outfile> is omitted, writes to stdout.  Needs a CUDA device: no CPU fallback."""
import argparse
import sys


def main(argv=None, engine=None):
    ap = argparse.ArgumentParser(prog="compressjs_flattened_b200", usage="%(prog)s -d|-z [infile] [outfile]", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-d", "--decompress", action="store_true", help="Decompress stdin to stdout")
    ap.add_argument("-z", "--compress", action="store_true", help="Compress stdin to stdout")
    ap.add_argument("-b", "--block", type=int, default=-1, metavar="<n>", help="Extract a single block, starting at <n> bits.")
    ap.add_argument("-t", dest="type", default="bzip2", metavar="<compressor>", help="Select compressor type (only bzip2 is on this path)")
    ap.add_argument("--multistream", action="store_true", help="with -d: decode every concatenated stream")
    ap.add_argument("--chunk-mb", type=int, default=64, help="input collected before its blocks are compressed")
    for lv in range(1, 10):
        ap.add_argument(f"-{lv}", dest=f"l{lv}", action="store_true",
                        help="Fastest/largest compression" if lv == 1 else "Slowest/smallest compression" if lv == 9 else argparse.SUPPRESS)
    ap.add_argument("infile", nargs="?", default=None)
    ap.add_argument("outfile", nargs="?", default=None)
    a = ap.parse_args(argv)
    if not a.decompress:
        a.compress = True
    if a.decompress and a.compress:
        print("Must specify either -d or -z.", file=sys.stderr)
        return 1
    if a.compress and a.block >= 0:
        print("--block can only be used with decompression", file=sys.stderr)
        return 1
    level = None
    for lv in range(1, 10):
        if getattr(a, f"l{lv}"):
            if level:
                print(f"Can't specify both -{level} and -{lv}", file=sys.stderr)
                return 1
            level = lv
    if level and a.decompress:
        print("Compression level has no effect when decompressing.", file=sys.stderr)
        return 1
    level = level or 7   # NPM/bin/compressjs:58
    if a.type.lower() not in ("bzip", "bzip2"):
        print(f"Unknown compressor on this path: {a.type} (only bzip2; SURVEY.md section 8)", file=sys.stderr)
        return 1
    if engine is None:
        from .bzip2 import Bzip2 as engine
    Bzip2 = engine
    src = open(a.infile, "rb") if a.infile else sys.stdin.buffer
    dst = open(a.outfile, "wb") if a.outfile else sys.stdout.buffer
    try:
        if a.decompress and a.block >= 0:
            dst.write(Bzip2.decompressBlock(src.read(), a.block))
        elif a.decompress:
            Bzip2.decompressStream(src, dst, a.multistream)
        else:
            Bzip2.compressStream(src, dst, level, chunk_bytes=a.chunk_mb << 20)
    finally:
        if a.infile:
            src.close()
        if a.outfile:
            dst.close()
        else:
            dst.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
