import ctypes
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, ORACLE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as entry
    entry.build()
    return entry


@pytest.fixture(scope="session")
def hostemu(built):
    """TEST-ONLY host emulation (warp width 1) of the device interpreter; never used by the product package."""
    from itsim_b200 import _lib
    return _lib.bind(ctypes.CDLL(built.HOSTEMU))


@pytest.fixture(scope="session")
def hostemu_fast(built):
    """The host emulation compiled like the in-place kernel: select()'s helper hops without the interpreter."""
    from itsim_b200 import _lib
    return _lib.bind(ctypes.CDLL(built.HOSTEMU.replace(".so", "_fast.so")))


@pytest.fixture(scope="session")
def golden_summaries():
    with open(os.path.join(GOLDEN, "netsimple_summaries.json")) as fh:
        return json.load(fh)


def load_golden_trace(name):
    return np.load(os.path.join(GOLDEN, name))["trace"]


def assert_trace_equal(got, ref, rtol=0.0):
    """Bit-exact on event order / kinds / programs / nodes / packet sizes AND on the timestamps (rtol = 0): every side
    computes them with the same IEEE operations in the same order -- no FMA contraction (-fmad=false /
    -ffp-contract=off), one shared logarithm (log_f64 in itsb_core.cuh == oracle/philox.py).  north_star allows 1e-9
    relative; the engine does not need the allowance."""
    assert len(got) == len(ref), f"trace length {len(got)} != {len(ref)}"
    for field in ("kind", "prog", "node", "aux"):
        neq = np.nonzero(got[field] != ref[field])[0]
        assert len(neq) == 0, f"first {field} mismatch at event {neq[0]}: got {got[neq[0]]}, ref {ref[neq[0]]}"
    if rtol == 0.0:
        neq = np.nonzero(got["t"] != ref["t"])[0]
        assert len(neq) == 0, f"first timestamp mismatch at event {neq[0]}: got {got['t'][neq[0]]!r}, ref {ref['t'][neq[0]]!r}"
    else:
        assert np.allclose(got["t"], ref["t"], rtol=rtol, atol=0.0)


def telemetry_matches_summary(tm, s):
    from itsim_b200 import _lib as L
    from itsim_b200.netsimple import STAT_QUERIES, STAT_RESOLVED, STAT_TIMEOUTS
    for k in range(1, L.NUM_KINDS):
        assert int(tm[L.TM_KIND0 + k]) == s["counts"][L.KIND_NAMES[k]], L.KIND_NAMES[k]
    assert int(tm[L.TM_PACKETS_SENT]) == s["packets_sent"]
    assert int(tm[L.TM_BYTES_SENT]) == s["bytes_sent"]
    assert int(tm[L.TM_PACKETS_RECEIVED]) == s["packets_received"]
    assert int(tm[L.TM_BYTES_RECEIVED]) == s["bytes_received"]
    assert int(tm[L.TM_USER0 + STAT_QUERIES]) == s["queries"]
    assert int(tm[L.TM_USER0 + STAT_RESOLVED]) == s["resolved"]
    assert int(tm[L.TM_USER0 + STAT_TIMEOUTS]) == s["timeouts"]
    assert int(tm[L.TM_DELIVERED]) + int(tm[L.TM_DROPPED]) == s["counts"]["DELIVER"]
