"""CPU tests of the boundary: libhtcn.so builds for sm_100a, loads, exports every symbol include/htcn.h
declares, the ctypes prototypes match the header's arity, and there is no silent fallback."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from hiertcn_b200.build import build
    return build()


def header_functions():
    src = open(os.path.join(ROOT, "include", "htcn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int32_t|int64_t|const char\*)\s+(htcn_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
    return out


def test_library_exports_every_declared_symbol(built):
    from hiertcn_b200 import _cabi as cabi
    lib = cabi.load(built)
    decl = header_functions()
    assert len(decl) >= 14
    for name in decl:
        assert hasattr(lib, name), "libhtcn.so does not export %s" % name
    assert lib.htcn_abi_version() == 6


def test_ctypes_prototypes_match_header(built):
    from hiertcn_b200 import _cabi as cabi
    decl = header_functions()
    bound = dict(cabi.SIGNATURES)
    bound.update({k: v[1] for k, v in cabi.PLAIN.items()})
    assert set(bound) == set(decl), set(bound) ^ set(decl)
    for name, n in decl.items():
        assert len(bound[name]) == n, (name, len(bound[name]), n)


def test_argument_validation_returns_error_not_crash(built):
    from hiertcn_b200 import _cabi as cabi
    cabi.load(built)
    with pytest.raises(cabi.HtcnError, match="bad"):
        cabi.call("htcn_prepare_wout", None, None, 0, None, 0, None)
    with pytest.raises(cabi.HtcnError):
        cabi.call("htcn_topk_merge", None, None, 1, 1, 1, None, None, None)


def test_missing_library_fails_loudly(tmp_path):
    from hiertcn_b200 import _cabi as cabi
    with pytest.raises(cabi.HtcnError, match="no CPU fallback"):
        cabi.load(str(tmp_path / "nope.so"))


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hiertcn_b200 import _cabi as cabi
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN
    with pytest.raises(cabi.HtcnError, match="no CPU fallback"):
        HierTCN(make_args(["--item_num", "50"])).build()


def test_product_never_imports_oracle():
    """only tests/, __graft_entry__.smoke and bench.py's cpu_baseline leg may touch oracle/"""
    pkg = os.path.join(ROOT, "hiertcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


def test_launch_table_is_consistent_with_the_signatures():
    """every entry of the per-call launch table names a bound function and is an int (bench.py's gpu_launches claim)"""
    from hiertcn_b200 import _cabi as cabi
    for name, n in cabi.LAUNCHES_PER_CALL.items():
        assert name in cabi.SIGNATURES, name
        assert isinstance(n, int) and n >= 0, (name, n)
