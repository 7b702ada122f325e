"""The C-ABI library loads without a GPU and exports every symbol include/cpm_b200.h declares."""
import ctypes as C

import pytest


def test_header_symbols_exported(cpm):
    lib = cpm.lib()
    names = cpm.declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/cpm_b200.h but not exported: {missing}"


def test_version_string(cpm):
    assert b"sm_100a" in cpm.lib().cpm_version()


def test_no_cpu_fallback(cpm):
    """Without a CUDA device context creation must fail loudly, never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(cpm.CpmError) as e:
        cpm.Context(0)
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)


def test_glibc_rand_reproduced(cpm):
    """cpm_rng_host_base_offsets == srand(seed); rand() of the C library
    (rng/mwc64xseedgenerator.cpp:56-64)."""
    libc = C.CDLL(None)
    for seed in (0, 1, 42, 0xFFFFFFFF):
        got = cpm.capi.rng_host_base_offsets(seed, 1000)
        libc.srand(C.c_uint(seed))
        want = [libc.rand() for _ in range(1000)]
        assert got[:, 0].tolist() == want
        assert (got[:, 1] == 0).all()


def test_host_header_symbols_exported():
    """libcpm_host.so (the reference-facing plugin layer) exports every CPMH_API function host_capi.h declares"""
    import importlib
    import re
    from conftest import PKG_NAME, ROOT
    host = importlib.import_module(PKG_NAME + ".host")
    text = (ROOT / PKG_NAME / "host" / "host_capi.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = sorted(set(re.findall(r"CPMH_API\s+[\w\s\*]+?\b(cpmh_\w+)\s*\(", text)))
    assert len(names) >= 40
    lib = host.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in host/host_capi.h but not exported: {missing}"
