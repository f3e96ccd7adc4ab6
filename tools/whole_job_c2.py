#!/usr/bin/env python
"""
BASELINE.json config[1] as written, once: network_simple x 100 000 replicas (seed 0, ids 0..99 999) advanced to the
example's own horizon of 12 h = 43 200 simulated seconds on ONE B200, as a whole job through the public API.

Checks: no replica fails (itsb_first_failed_replica == -1: the arena, the event heap and the process table hold for every
replica over the whole horizon), the job's counters obey the conservation laws, and the replicas that have a golden
summary at this horizon (ids 0, 1, 2, 3, 63 and 99 999: oracle/make_golden.py --full) equal the oracle counter for
counter.  Prints one JSON line (kept under profiles/).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=100000)
    ap.add_argument("--horizon", type=float, default=43200.0)
    ap.add_argument("--slice", type=float, default=3600.0)
    ap.add_argument("--arena-nodes", type=int, default=16384)
    args = ap.parse_args()
    import numpy as np
    from itsim_b200 import netsimple, _lib as L
    from itsim_b200.engine import summarize
    from conftest import telemetry_matches_summary
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "netsimple_summaries.json")))

    t0 = time.perf_counter()
    sim = netsimple.build()
    sim.caps["arena_nodes"] = args.arena_nodes
    eng = sim.run_replicas(args.replicas, seed=0)
    t_cold = time.perf_counter() - t0
    events, device_ms, t = 0, 0.0, 0.0
    per_slice = []
    while t < args.horizon - 1e-9:
        dt = min(args.slice, args.horizon - t)
        ev = eng.run(dt)
        events += ev
        device_ms += eng.last_run_ms()
        per_slice.append({"until": t + dt, "events": ev, "ms": eng.last_run_ms()})
        t += dt
    tm = eng.telemetry()
    wall = time.perf_counter() - t0
    failed = int(eng._lib.itsb_first_failed_replica(eng._h))
    checked = {}
    for r in (0, 1, 2, 3, 63, 99999):
        key = f"0:{r}:{int(args.horizon)}"
        if r < args.replicas and key in golden:
            mine = eng.replica_telemetry(r).copy()
            mine[L.TM_KIND0 + L.EV_STOP] -= len(per_slice) - 1        # one STOP per run() call; the oracle made one call
            telemetry_matches_summary(mine, golden[key])                 # raises on the first mismatch
            checked[str(r)] = golden[key]["events"]
    s = summarize(tm)
    k0 = L.TM_KIND0
    laws = {
        "events_equal_sum_of_kinds": int(tm[:9].sum()) == events + 0 * s["events"],
        "delivered_plus_dropped_equal_deliver_events": int(tm[L.TM_DELIVERED]) + int(tm[L.TM_DROPPED]) == int(tm[k0 + L.EV_DELIVER]),
        "tx_begin_equal_packets_sent": int(tm[k0 + L.EV_TX_BEGIN]) == int(tm[L.TM_PACKETS_SENT]),
        "every_replica_at_the_horizon": bool(np.all(eng.now() == args.horizon)),
        "stop_events": int(tm[k0 + L.EV_STOP]) == args.replicas * len(per_slice),
    }
    line = {
        "job": f"BASELINE.json config[1]: network_simple x {args.replicas} replicas x {args.horizon:g} simulated s, 1 GPU",
        "kernel": eng.kernel_name(), "events": events, "device_s": device_ms * 1e-3,
        "events_per_s_device": events / (device_ms * 1e-3), "wall_s": wall, "cold_s": t_cold,
        "events_per_s_whole_job": events / wall, "first_failed_replica": failed,
        "replicas_equal_to_the_oracle_at_the_horizon": checked, "laws": laws,
        "telemetry": {k: v for k, v in s.items() if k != "latency_hist"}, "slices": per_slice,
        "arena_nodes": args.arena_nodes, "state_bytes": eng.state_bytes(),
    }
    assert failed == -1 and all(laws.values()), (failed, laws)
    print(json.dumps(line))
    eng.close()


if __name__ == "__main__":
    main()
