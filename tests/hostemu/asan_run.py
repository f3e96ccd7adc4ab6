"""Runs every model family through the device source compiled at warp width 1 with AddressSanitizer and
UndefinedBehaviorSanitizer (tests/test_sanitizers.py builds the library and preloads the runtimes).  compute-sanitizer
is not available on the GPU pool; this is the memory-safety check of the same source, table layouts and arena
bookkeeping the CUDA build uses.  argv[1] = the sanitized library."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from itsim_b200 import _lib as L, machine_models, netsimple  # noqa: E402

lib = L.bind(ctypes.CDLL(sys.argv[1]))


def run(name, sim, n, horizon, **caps):
    sim.caps.update(caps)
    sim._lib = lib
    eng = sim.run_replicas(n, seed=3)
    events = eng.run(horizon)
    print(f"{name}: {n} x {horizon} s, {events} events", flush=True)
    eng.close()


run("network_simple", netsimple.build(), 2, 900.0)
run("network_simple traced", netsimple.build(trace_records=60000), 1, 600.0)
run("network_simple tiny pool", netsimple.build(), 1, 900.0, max_events=16)
run("network_simple plain loops", netsimple.build(fused=False), 1, 600.0)
run("full logging", netsimple.build(connection_records=20000), 1, 900.0, net_event_records=4096)
run("full logging, plain receive loops", netsimple.build(connection_records=20000, fused=False), 1, 900.0,
    net_event_records=4096)
run("full logging, window 1, logs overflowing", netsimple.build(connection_records=200, connection_window=1), 1, 900.0,
    net_event_records=16)
run("malware", netsimple.build_malware(), 2, 600.0)
run("malware logged", netsimple.build_malware(connection_records=5000), 1, 600.0)
run("backdoor", netsimple.build_backdoor(), 1, 7200.0)
for name in ("packet_transfer", "broadcast", "thread_lifecycle", "process_lifecycle", "dhcp_exchange", "link_delay"):
    run(name, getattr(machine_models, name)(), 2, 130.0)
run("large LAN", machine_models.large_lan(), 1, 0.25)
print("clean")
