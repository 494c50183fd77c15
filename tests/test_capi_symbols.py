"""The C-ABI library loads on a CPU-only box and exports every symbol include/bos_b200.h declares;
compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import pytest

from prb_project_bearing_only_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "bos_b200.h")).read()
    declared = re.findall(r"BOS_API\s+[\w\s\*]+?\b(bos_\w+)\s*\(", hdr)
    assert len(declared) >= 40
    assert sorted(set(declared)) == sorted(capi.SYMBOLS)
    for s in declared:
        assert hasattr(built_lib, s), s


def test_no_cpu_fallback_without_a_device(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.BosError) as e:
        capi.Context()
    assert e.value.code == capi.ERR_CUDA


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "prb_project_bearing_only_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                for needle in ("bos_oracle", "oracle/", "from oracle", "import oracle", "oracle.oracle"):
                    assert needle not in txt, (os.path.join(dirpath, f), needle)
