"""The C-ABI library loads and exports exactly what include/vast_b200.h declares (no compute calls: CPU ok)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vast_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return re.findall(r"VAST_API\s+[\w\s\*]+?\b(vast_\w+)\s*\(", src)


def test_header_symbols_exported_and_bound():
    from vast_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    names = _declared()
    assert len(names) >= 25 and len(set(names)) == len(names)
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported by the library"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert L.vast_version() == 100
    assert isinstance(L.vast_last_error_string(), bytes)


def _ctype_of(decl: str):
    """ctypes type a C parameter / return declaration of include/vast_b200.h maps to (pointers -> c_void_p)."""
    import ctypes as C
    d = re.sub(r"/\*.*?\*/", "", decl, flags=re.S).strip()
    if "*" in d:
        return C.c_char_p if re.match(r"const\s+char\s*\*\s*$", d) else C.c_void_p
    words = [w for w in re.split(r"\s+", d) if w and w != "const"]
    if len(words) > 1 and re.match(r"[A-Za-z_]\w*$", words[-1]) and words[-1] not in ("int", "float", "double", "size_t"):
        words = words[:-1]                                 # drop the parameter name
    ty = " ".join(words)
    table = {"int": C.c_int, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "int32_t": C.c_int32, "uint32_t": C.c_uint32,
             "float": C.c_float, "double": C.c_double, "size_t": C.c_size_t, "vast_stream_t": C.c_void_p}
    assert ty in table, f"unmapped C type {ty!r} in {decl!r}"
    return table[ty]


def test_argument_types_match_header():
    """Every ctypes signature in vast_b200/_lib.py has the return type and the argument TYPES (not just the count)
    that include/vast_b200.h declares, parameter by parameter."""
    import ctypes as C
    from vast_b200 import _lib
    src = open(os.path.join(ROOT, "include", "vast_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    seen = 0
    for m in re.finditer(r"VAST_API\s+([\w\s\*]+?)\b(vast_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        want = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",") if a.strip()]
        res, got = _lib.SIGNATURES[name]
        assert len(want) == len(got), (name, len(want), len(got))
        for i, (w, g) in enumerate(zip(want, got)):
            assert C.sizeof(w) == C.sizeof(g) and (w is g or {w, g} <= {C.c_int, C.c_int32}), (name, i, w, g)
        want_res = C.c_char_p if re.match(r"const\s+char\s*\*$", ret) else _ctype_of(ret + " x")
        assert want_res is res or {want_res, res} <= {C.c_int, C.c_int32}, (name, ret, res)
        seen += 1
    assert seen == len(_lib.SIGNATURES)


def test_no_cpu_fallback():
    """Ops refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    from vast_b200 import ops
    with pytest.raises(RuntimeError):
        ops.l2norm(torch.zeros(4, 8))
    with pytest.raises(RuntimeError):
        ops.gemm_nt(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))


def test_product_does_not_import_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "vast_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
