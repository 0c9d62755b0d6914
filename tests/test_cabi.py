"""CPU-side checks of the C-ABI boundary: the library builds/loads without a GPU and exports every symbol
include/occdepth_b200.h declares; host-side argument validation answers without touching a device."""
import ctypes as C
import os
import re

from occdepth_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "occdepth_b200.h")).read()
    return sorted(set(re.findall(r"\b(occd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    _build.build()
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), n
        assert n in _lib.SYMBOLS, "ctypes binding missing for " + n
    assert L.occd_abi_version() == 2


def test_argument_validation_without_gpu():
    L = _lib.lib()
    p = _lib.SfaParams()
    assert L.occd_sfa_lift_fwd(C.byref(p), None) != 0
    assert b"n_scales" in L.occd_last_error()
    d = _lib.ConvDesc()
    h = C.c_void_p()
    assert L.occd_conv_plan_create(C.byref(d), C.byref(h)) != 0
    assert L.occd_planar_to_cl(None, None, 0, 1, 1, 1, 1, None) != 0


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "occdepth_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)


def _prototypes():
    """name -> list of parameter C types, parsed from include/occdepth_b200.h (comments stripped)"""
    src = open(os.path.join(ROOT, "include", "occdepth_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\s*\*)\s*(occd_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = " ".join(m.group(2).split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[m.group(1)] = plist
    return protos


def _ctype_class(p):
    """coarse class of a C parameter declaration"""
    if "*" in p:
        return "ptr"
    base = " ".join(p.split()[:-1])       # drop the parameter name
    return {"int": "int", "long long": "ll", "float": "float", "double": "double"}.get(base.replace("const ", ""), base)


def test_ctypes_bindings_match_the_header_prototypes():
    """arity and parameter classes of every ctypes binding (occdepth_b200/_lib.py SYMBOLS) against the header: a
    drifted binding would silently shift arguments on the way into the library"""
    protos = _prototypes()
    assert set(protos) == set(_declared())
    cls = {C.c_void_p: "ptr", C.c_int: "int", C.c_longlong: "ll", C.c_float: "float", C.c_double: "double"}
    for name, params in protos.items():
        _, argtypes = _lib.SYMBOLS[name]
        assert len(argtypes) == len(params), (name, len(argtypes), params)
        for a, p in zip(argtypes, params):
            want = _ctype_class(p)
            got = cls.get(a, "ptr")            # POINTER(struct) / POINTER(c_int) / c_char_p are pointers
            assert got == want, (name, p, a)
