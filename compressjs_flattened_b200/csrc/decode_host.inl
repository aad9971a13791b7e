// decode_host.inl -- host side of the decompress path (included by bz2b200.cu)
