"""The oracle against its pins: the reference's own tests (when /root/reference is present) and the committed
golden vectors."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ORACLE, assert_trace_equal, load_golden_trace

REFERENCE = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "itsim")), reason="reference tree not on this box")
def test_reference_suite_passes_on_the_shim():
    """151 hot-path tests of the reference, unmodified, on the unmodified reference itsim + oracle/greensim."""
    proc = subprocess.run([sys.executable, os.path.join(ORACLE, "run_reference_tests.py")], capture_output=True,
                          text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-4000:]
    assert "failing on the shim: 0" in proc.stdout


def test_oracle_reproduces_golden_trace():
    import netsimple
    ref = load_golden_trace("netsimple_trace_s7_r3_900s.npz")
    model = netsimple.run_replica(7, 3, 900.0)
    got = model.tracer.as_array()
    model.close()
    assert_trace_equal(got, ref, rtol=0.0)


def test_oracle_summary_counts(golden_summaries):
    import netsimple
    s = golden_summaries["7:3:900"]
    model = netsimple.run_replica(7, 3, 900.0, trace=False)
    out = model.tracer.summary()
    model.close()
    assert out["counts"] == s["counts"]
    assert model.rng.draws == s["draws"]
    assert model.stats["bytes_sent"] == s["bytes_sent"]


def test_event_definition_excludes_helper_hops():
    """select()/balk helpers are engine plumbing: hidden hops, not simulated events."""
    import greensim
    import evtrace
    sim = greensim.Simulator()
    tracer = evtrace.Tracer(sim)
    sig_a, sig_b = greensim.Signal().turn_off(), greensim.Signal().turn_off()
    log = []

    def waiter():
        try:
            greensim.select(sig_a, sig_b, timeout=5.0)
        except greensim.Timeout:
            log.append(greensim.now())

    evtrace.Tracer.label(sim.add(waiter), evtrace.ROLE_MODEL, 1, 0)
    sim.run(10.0)
    assert log == [5.0]
    kinds = [r[1] for r in tracer.records]
    assert kinds == [evtrace.PROC_START, evtrace.THROW, evtrace.STOP]
    assert tracer.hidden > 0
