"""CPU: the C-ABI library loads, exports every symbol include/vieo_b200.h declares, and fails loudly
(no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "vieo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vieo_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import vieo_slam_b200.api as api
    L = api.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/vieo_b200.h but not exported"
    assert b"sm_100a" in L.vieo_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vieo_slam_b200.api as api
    with pytest.raises(api.VieoError, match="no CUDA device|CUDA"):
        api.ORBextractor(1200, 1.2, 8, 20, 7, 752, 480)
    import numpy as np
    with pytest.raises(api.VieoError):
        api.ORBmatcher().knnMatch2(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))


def test_product_does_not_import_oracle():
    # the oracle is test infrastructure: nothing under vieo_slam_b200/ may reference it
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "vieo_slam_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                s = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle_lib|liboracle|oracle/|orc_", s):
                    bad.append(f)
    assert not bad, bad
