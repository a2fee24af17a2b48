"""CPU-only checks of the C ABI library: it loads, exports every symbol include/slrbs_b200.h
declares, and fails loudly (no CPU fallback) when no GPU is present."""
import os
import re

import pytest

import slrbs_b200
from slrbs_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _built():
    return os.path.exists(capi.lib_path())


def test_header_symbols_all_exported():
    import __graft_entry__ as ge
    if not _built():
        ge.build()
    hdr = open(os.path.join(ROOT, "include", "slrbs_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(slrbs_[a-z_0-9]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    L = slrbs_b200.load_library()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(declared) == sorted(capi.SYMBOLS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(slrbs_b200.SlrbsError):
        slrbs_b200.Context(1, 4, 0, 4)


def test_product_does_not_reference_oracle():
    bad = []
    for base in ("slrbs_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"oracle_lib|libslrbs_oracle|slrbs_oracle\.h|\borc_[a-z_]+\(|import oracle|from oracle", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
