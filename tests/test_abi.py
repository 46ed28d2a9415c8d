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


def test_argument_counts_match_header():
    from vast_b200 import _lib
    src = open(os.path.join(ROOT, "include", "vast_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for m in re.finditer(r"VAST_API\s+[\w\s\*]+?\b(vast_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))


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
