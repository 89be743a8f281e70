"""CPU: the oracle reproduces its own committed outputs (tests/golden/oracle_goldens.npz, written by
tests/golden/gen_oracle_goldens.py).  These fixtures pin the ORACLE against accidental edits — they are not an
independent reference; correctness is pinned by the cv2 goldens, the naive restatements and the numeric checks of the
other test_oracle_* files.  Integers exact, floating point 1e-9 relative to the array's largest magnitude."""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_committed_outputs():
    spec = importlib.util.spec_from_file_location("gen_oracle_goldens", os.path.join(HERE, "golden", "gen_oracle_goldens.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    now = gen.compute()
    ref = np.load(os.path.join(HERE, "golden", "oracle_goldens.npz"))
    assert sorted(now) == sorted(ref.files)
    for k in ref.files:
        a, b = now[k], ref[k]
        assert a.shape == b.shape, k
        if b.dtype.kind in "iub":
            assert np.array_equal(a, b), (k, np.nonzero(a.ravel() != b.ravel())[0][:8])
        else:
            na, nb = np.isnan(a), np.isnan(b)
            assert np.array_equal(na, nb), k
            scale = max(np.abs(b[~nb]).max(), 1e-300) if (~nb).any() else 1.0
            assert np.abs(a[~na] - b[~nb]).max() <= 1e-9 * scale, (k, np.abs(a[~na] - b[~nb]).max(), scale)
