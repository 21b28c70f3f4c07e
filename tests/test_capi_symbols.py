"""CPU: the C-ABI library loads and exports every symbol include/folve_b200.h declares."""
import ctypes
import os

import pytest

from folve_b200 import capi


def test_library_exports_every_declared_symbol():
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libfolve_b200.so not built (run `make`)")
    L = ctypes.CDLL(capi.LIB_PATH)
    names = capi.declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert L.fcv_abi_version() == 2


def test_binding_covers_header():
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libfolve_b200.so not built")
    L = capi.lib()
    for n in capi.declared_symbols():
        assert getattr(L, n).argtypes is not None, n


def test_no_cpu_fallback_without_gpu():
    """Without a usable device the product path fails loudly instead of computing on the CPU."""
    if not os.path.exists(capi.LIB_PATH):
        pytest.skip("libfolve_b200.so not built")
    L = capi.lib()
    if L.fcv_device_count() > 0:
        pytest.skip("a GPU is present")
    f = capi.Filter(1, 1, 100, 64)
    f.add(0, 0, [1.0], 0)
    with pytest.raises(capi.FcvError):
        f.commit(0)
