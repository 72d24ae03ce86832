"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/rsrl_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rsrl_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return abi.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "rsrl_b200.h")).read()
    declared = {n for n in re.findall(r"\b(rsrl_[a-z0-9_]+)\s*\(", hdr) if not n.endswith("_t")}
    assert declared == set(abi.SYMBOLS), declared ^ set(abi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_matches_header(lib):
    cfg = abi.Config()
    assert lib.rsrl_config_default(C.byref(cfg)) == 0
    assert cfg.struct_size == C.sizeof(abi.Config)
    want = abi.default_config()
    assert bytes(cfg) == bytes(want)
    assert lib.rsrl_version() == 1


def test_config_dims_and_validation(lib):
    from rsrl_b200.engine import config_dims
    assert config_dims(abi.default_config()) == (2, 3, 36)
    assert config_dims(abi.default_config(domain=abi.ACROBOT, basis_order=7)) == (4, 3, 4096)
    bad = abi.default_config(basis_order=0)
    d = C.c_int32()
    assert lib.rsrl_config_dims(C.byref(bad), C.byref(d), None, None) == abi.EINVAL
    assert b"basis_order" in lib.rsrl_last_error()
    bad = abi.default_config()
    bad.struct_size = 8
    assert lib.rsrl_config_dims(C.byref(bad), C.byref(d), None, None) == abi.EINVAL
    bad = abi.default_config(algo=abi.TD_LAMBDA, policy=abi.GREEDY)
    assert lib.rsrl_config_dims(C.byref(bad), C.byref(d), None, None) == abi.EINVAL


def test_domain_info_matches_oracle(lib, oracle):
    from rsrl_b200.engine import domain_info
    for dom in (abi.MOUNTAIN_CAR, abi.CART_POLE, abi.ACROBOT):
        D, A, lo, hi, start = domain_info(dom)
        assert (D, A) == oracle.domain_dims(dom)
        olo, ohi = oracle.domain_limits(dom)
        assert (lo == olo).all() and (hi == ohi).all() and (start == oracle.domain_default(dom)).all()


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product refuses to compute (it never routes through oracle/)."""
    if lib.rsrl_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    cfg = abi.default_config()
    assert lib.rsrl_engine_create(C.byref(cfg), C.byref(h)) == abi.ENODEVICE
    assert b"no CPU fallback" in lib.rsrl_last_error()
    s = np.zeros((1, 2))
    r, t = np.zeros(1), np.zeros(1, dtype=np.uint8)
    a = np.zeros(1, dtype=np.int32)
    assert lib.rsrl_domain_step(0, 1, abi.dp(s), abi.ip(a), abi.dp(r), abi.u8p(t)) == abi.ENODEVICE


def test_product_never_imports_oracle():
    """The product path must not import / include / link anything under oracle/."""
    pat = re.compile(r"(^\s*(import|from)\s+\S*oracle|#\s*include\s*[\"<][^\">]*oracle|dlopen\([^)]*oracle|CDLL\([^)]*oracle)", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rsrl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), os.path.join(dirpath, f)
