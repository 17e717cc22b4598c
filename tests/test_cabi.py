"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/ruart_b200.h declares (no compute calls here)."""
import ctypes

from ruart_b200 import _lib, build


def test_library_builds_and_exports_all_declared_symbols():
    path = build.build()
    L = ctypes.CDLL(path)
    names = _lib.declared_symbols()
    assert len(names) >= 7
    for n in names:
        assert hasattr(L, n), "missing export %s" % n
    for n in _lib._SIGNATURES:
        assert n in names, "binding %s is not declared in the header" % n
    for n in names:
        assert n in _lib._SIGNATURES, "declared symbol %s has no ctypes binding" % n


def test_version_and_error_string():
    L = _lib.lib()
    assert L.ruart_version() >= 100
    assert isinstance(L.ruart_last_error(), bytes)


def test_bad_arguments_are_reported_without_a_gpu():
    L = _lib.lib()
    # Kp not a multiple of 64 -> RUART_ERR_ARG before any CUDA call
    rc = L.ruart_gemm_bf16(None, 8, 1, None, 8, 1, 4, 4, 10, 1, 0, None, None, 0, None, 0, None, 0,
                           1, 0, 0, None, 0, None)
    assert rc == 2
    assert b"bad argument" in L.ruart_last_error()


def test_ctypes_signatures_have_the_arity_the_header_declares():
    # every prototype in include/ruart_b200.h: the number of parameters equals the length of the
    # ctypes argtypes list (catches ABI drift between the header, the library and the binding)
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "ruart_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = re.findall(r"RUART_API\s+[\w\s\*]+?\b(ruart_\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
    assert len(protos) == len(_lib.declared_symbols())
    for name, args in protos:
        args = args.strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n == len(_lib._SIGNATURES[name]), "%s: header %d parameters, binding %d" % (name, n, len(_lib._SIGNATURES[name]))
