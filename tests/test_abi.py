"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/kmer_id.h declares, and compute calls fail loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import helpers as H


def _declared():
    src = open(os.path.join(H.ROOT, "include", "kmer_id.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kid_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import kmer_id_b200 as kid
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(kid.lib, n), f"{n} declared in include/kmer_id.h but not exported"
    assert sorted(kid._SIGS) == names, "python binding and header disagree"


def test_no_cpu_fallback():
    import kmer_id_b200 as kid
    if kid.device_count() > 0:
        pytest.skip("GPU present")
    parent = np.ones(10, np.int32)
    with pytest.raises(kid.KidError) as e:
        kid.Database(np.array([1], np.uint64), np.array([2], np.uint32), parent)
    assert e.value.code == -2  # KID_ECUDA


def test_product_never_touches_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(H.ROOT, "kmer_id_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"oracle|kid_oracle|liboracle|kor_", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files reference the oracle: {bad}"
