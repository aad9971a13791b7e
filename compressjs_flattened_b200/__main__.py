"""Command line in the spirit of the reference's bin/compressjs (NPM/bin/compressjs:60-180), bzip2 only:

  python -m compressjs_flattened_b200 [-z | -d] [-1 .. -9] [-t bzip2] [-o OUT] [IN]

-z compresses (stream flavour: the input is read in chunks, compressStream), -d decompresses (whole file, multistream).
IN / OUT default to stdin / stdout.  Needs a CUDA device: there is no CPU fallback."""
import argparse
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(prog="compressjs_flattened_b200", description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    g = ap.add_mutually_exclusive_group()
    g.add_argument("-z", "--compress", action="store_true", help="compress (default)")
    g.add_argument("-d", "--decompress", action="store_true", help="decompress")
    ap.add_argument("-t", "--type", default="bzip2", help="codec; only bzip2 is on this path")
    ap.add_argument("-o", "--output", default=None)
    ap.add_argument("--chunk-mb", type=int, default=64, help="input chunk of the streaming compressor")
    for lv in range(1, 10):
        ap.add_argument(f"-{lv}", dest="level", action="store_const", const=lv, help=argparse.SUPPRESS)
    ap.add_argument("input", nargs="?", default=None)
    a = ap.parse_args(argv)
    if a.type.lower() != "bzip2":
        ap.error("only -t bzip2 is implemented on this path (SURVEY.md section 8)")
    from .bzip2 import Bzip2
    src = open(a.input, "rb") if a.input else sys.stdin.buffer
    dst = open(a.output, "wb") if a.output else sys.stdout.buffer
    try:
        if a.decompress:
            dst.write(Bzip2.decompressFile(src.read(), None, True))
        else:
            Bzip2.compressStream(src, dst, a.level or 9, chunk_bytes=a.chunk_mb << 20)
    finally:
        if a.input:
            src.close()
        if a.output:
            dst.close()
        else:
            dst.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
