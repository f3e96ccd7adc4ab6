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


# ---- the reference's own example files, executed unmodified (oracle/ref_example.py) --------------------------------
needs_example = pytest.mark.skipif(
    not os.path.isfile(os.path.join(REFERENCE, "examples", "network_simple", "network_simple.py")),
    reason="reference tree not on this box")


def _assert_same_run(ref, ours):
    """Bit for bit: every column of the trace (timestamps included: the same float operations on the same draws),
    the connection records the override Socket logs, the draws consumed, the final clock."""
    assert ref["summary"]["events"] == ours["summary"]["events"]
    assert ref["summary"]["counts"] == ours["summary"]["counts"]
    assert_trace_equal(ref["trace"], ours["trace"], rtol=0.0)
    assert ref["connections"] == ours["connections"]
    assert ref["draws"] == ours["draws"] and ref["now"] == ours["now"]


@needs_example
@pytest.mark.parametrize("seed,replica,duration", [(0, 0, 900.0), (7, 3, 700.0), (42, 17, 650.0), (123456789012, 5, 650.0)])
def test_restatement_equals_the_reference_example_run_unmodified(seed, replica, duration):
    """examples/network_simple/network_simple.py + network_simple_overrides.py, run as __main__ where they lie, against
    oracle/netsimple.py on the same Philox stream: this is what pins the flagship workload's oracle (SURVEY rows
    a16-a17) to the reference's own code instead of to a restatement of it."""
    import ref_example
    both = ref_example.compare_with_restatement("network_simple.py", seed, replica, duration)
    _assert_same_run(both["reference"], both["restatement"])
    assert len(both["reference"]["connections"]) > 0


@needs_example
@pytest.mark.parametrize("seed,replica,duration", [(0, 0, 500.0), (2, 41, 450.0), (9, 1000000, 400.0)])
def test_malware_restatement_equals_the_reference_script_run_unmodified(seed, replica, duration):
    """examples/network_simple/malware_simple.py (module-level script: init, Internet, infected = next(distribution(..)),
    route, command_and_control, sim.run) against oracle/netsimple.MalwareModel (row a18)."""
    import ref_example
    both = ref_example.compare_with_restatement("malware_simple.py", seed, replica, duration)
    _assert_same_run(both["reference"], both["restatement"])
    kinds = both["reference"]["trace"]["prog"]
    assert (kinds == 14).sum() > 0 and (kinds == 15).sum() > 0     # exfiltrate and route really ran


@needs_example
def test_golden_traces_are_what_the_reference_example_produces_now():
    """The committed golden vectors the GPU tests compare against, regenerated from the reference's files here."""
    import ref_example
    out = ref_example.run_network_simple(7, 3, 900.0)
    assert_trace_equal(out["trace"], load_golden_trace("netsimple_trace_s7_r3_900s.npz"), rtol=0.0)
    out = ref_example.run_malware_simple(0, 0, 900.0)
    assert_trace_equal(out["trace"], load_golden_trace("malware_trace_s0_r0_900s.npz"), rtol=0.0)


@pytest.mark.skipif(not os.path.isdir("/root/reference/itsim"), reason="reference tree not on this box")
def test_every_reference_citation_points_at_an_existing_line():
    """`file.py:LINE` citations in the docs, the ABI header and the sources: the file exists in the reference and is
    that long (tools/check_citations.py)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "check_citations.py")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
