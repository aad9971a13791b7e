"""The C-ABI library loads and exports every symbol include/bz2b200.h declares; without a GPU every
entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT

SO = os.path.join(ROOT, "compressjs_flattened_b200", "libbz2b200.so")
HDR = os.path.join(ROOT, "include", "bz2b200.h")


def _declared():
    text = open(HDR).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bz2b200_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(SO):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(SO)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bz2b200.h but not exported"


def test_product_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", SO], capture_output=True, text=True).stdout
    assert "orc_" not in out
    ldd = subprocess.run(["ldd", SO], capture_output=True, text=True).stdout
    assert "liboracle" not in ldd


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Bzip2Engine(0)
