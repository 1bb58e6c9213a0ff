"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the public
header declares, the ctypes mirror covers the same set, and the product path fails loudly (no CPU
fallback) when no GPU is present."""
import ctypes
import re
import subprocess

import pytest

import lqcd_b200 as q
from lqcd_b200 import _lib


def header_symbols():
    txt = _lib.HEADER.read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lqcd_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_loads():
    import __graft_entry__ as g
    g.build()
    lib = _lib.load()
    assert lib.lqcd_abi_version() == 2


def test_every_header_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 30
    lib = ctypes.CDLL(str(_lib.SO_PATH))
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lqcd_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, set(syms) ^ set(_lib.SIGNATURES)


def test_sm100a_only_cubin():
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.SO_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(q.LqcdError) as e:
        q.get_context((4, 4, 4, 4))
    assert e.value.code == _lib.ERR_NOGPU


def test_bad_arguments_are_reported_not_crashed():
    lib = _lib.load()
    out = ctypes.c_void_p()
    d = (ctypes.c_int * 4)(4, 4, 4, 4)
    p = (ctypes.c_int * 4)(1, 1, 1, 3)
    st = lib.lqcd_ctx_create(d, p, 0, 0, ctypes.byref(out))
    assert st != 0 and lib.lqcd_last_error(None)


def test_null_handles_are_argument_errors_everywhere():
    """every entry point that takes a context must answer a NULL context (and NULL operands) with LQCD_ERR_ARG and a message --
    never a crash -- so a binding mistake on the Julia side surfaces as error(...) (SURVEY.md 8b "Errors")"""
    lib = _lib.load()
    skip = {"lqcd_last_error", "lqcd_abi_version", "lqcd_decompose", "lqcd_io_read_gauge", "lqcd_io_write_gauge", "lqcd_ctx_create", "lqcd_ctx_destroy"}
    for name, (res, args) in sorted(_lib.SIGNATURES.items()):
        if name in skip:
            continue
        vals = []
        for a in args:
            if a in (ctypes.c_int, ctypes.c_uint64, ctypes.c_size_t):
                vals.append(a(1))
            elif a is ctypes.c_double:
                vals.append(a(1.0))
            else:
                vals.append(None)                      # every pointer NULL, the context included
        st = getattr(lib, name)(*vals)
        if name in ("lqcd_fermion_free", "lqcd_host_unregister"):       # releasing nothing is fine (finalizers)
            assert st in (0, _lib.ERR_ARG), (name, st)
            continue
        assert st == _lib.ERR_ARG, (name, st)
        assert lib.lqcd_last_error(None), name
    assert lib.lqcd_ctx_destroy(None) == 0              # destroying nothing is fine (finalizers)
    d = (ctypes.c_int * 4)(4, 4, 4, 4)
    assert lib.lqcd_io_read_gauge(None, 0, d, 3, None, 0) == _lib.ERR_ARG
    assert lib.lqcd_io_write_gauge(b"/nonexistent/dir/x.ildg", 0, d, 3, None, 0) == _lib.ERR_ARG
