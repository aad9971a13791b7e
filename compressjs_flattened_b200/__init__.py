"""B200-native bzip2 block compressor / decompressor behind compressjs' Bzip2 API."""
from .bzip2 import Bzip2, Bzip2Engine, Bzip2Error, Err  # noqa: F401
