"""
ORACLE / TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/ from the oracle (run in the build container).

The reference holds no seeded traces ("no test fixes an RNG seed"), so the stochastic golden
vectors are the oracle's own output under the injected Philox stream:

* netsimple_trace_s{seed}_r{replica}_{T}s.npz   full event trace (t, kind, prog, node, aux) of one replica
* netsimple_summaries.json                      per (seed, replica, horizon): event counts per kind, model counters,
                                                draws consumed, final clock, digest of the integer trace columns

* machine_<scenario>*.npz / machine_summaries.json   (--machine) traces of the reference's own integration-test
                                                bodies for the shipped-itsim primitives, oracle/machine_scenarios.py

Usage: python oracle/make_golden.py [--full]     (--full adds the 43 200 s = 12 h horizon of config C1: ~7 min)
       python oracle/make_golden.py --machine
       python oracle/make_golden.py --frontend   (tests/frontend_bodies.py run on the reference's itsim)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import evtrace  # noqa: E402
import netsimple  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def summary_of(model, horizon):
    arr = model.tracer.as_array()
    out = model.tracer.summary()
    out.update(model.stats)
    out.update({"draws": model.rng.draws, "now": model.sim.now(), "horizon": horizon,
                "connections": len(model.connections), "digest": evtrace.trace_digest(arr),
                "t_sum": float(np.sum(arr["t"]))})
    return out, arr


def machine_golden() -> None:
    """Golden traces of the shipped-itsim scenarios: the reference's own integration-test bodies (needs
    /root/reference)."""
    import machine_scenarios as ms
    out = {}
    jobs = [("thread_lifecycle", {}), ("process_lifecycle", {}),
            ("packet_transfer", {"seed": 0, "replica": 0}), ("packet_transfer", {"seed": 5, "replica": 77}),
            ("broadcast", {"seed": 0, "replica": 0}), ("broadcast", {"seed": 3, "replica": 11, "duration": 60.0}),
            ("dhcp_exchange", {"seed": 0, "replica": 0}), ("dhcp_exchange", {"seed": 4, "replica": 9, "duration": 100.0}),
            # config C4: small instance kept as a full trace, the 16 x 64 instance as counts + digest (~45 s of oracle)
            ("large_lan", {"seed": 1, "replica": 2, "duration": 20.0, "subnets": 4, "per_subnet": 16}),
            ("large_lan", {"seed": 0, "replica": 0, "duration": 8.0, "subnets": 16, "per_subnet": 64})]
    for name, kwargs in jobs:
        arr, summary = ms.SCENARIOS[name](**kwargs)
        tag = name + "".join(f"_{k[0]}{int(v)}" for k, v in kwargs.items())
        summary["t_digest"] = evtrace.trace_digest(arr)
        summary["t_last"] = float(arr["t"][-1]) if len(arr) else 0.0
        if len(arr) < 60000:
            np.savez_compressed(os.path.join(GOLDEN, f"machine_{tag}.npz"), trace=arr)
        out[tag] = summary
        print("machine", tag, summary["events"], flush=True)
    json.dump(out, open(os.path.join(GOLDEN, "machine_summaries.json"), "w"), indent=1, sort_keys=True)


def frontend_golden() -> None:
    """tests/frontend_bodies.py executed on the reference's itsim (needs /root/reference): what the engine must
    reproduce when the front-end compiles the same file."""
    import machine_scenarios as ms
    out = {}
    jobs = [("ping_pong", 0, 0, 10.0), ("ping_pong", 9, 4, 10.0), ("chatter", 1, 5, 30.0), ("threads", 0, 0, None),
            ("family", 0, 0, 4000.0), ("lan", 2, 3, 15.0), ("echo", 0, 0, 10.0), ("echo", 6, 21, 10.0),
            ("daemon", 0, 0, 10.0), ("daemon", 11, 2, 10.0)]
    for name, seed, replica, duration in jobs:
        arr, summary = ms.frontend_scenario(name, seed, replica, duration)
        tag = f"{name}_s{seed}_r{replica}"
        np.savez_compressed(os.path.join(GOLDEN, f"frontend_{tag}.npz"), trace=arr)
        out[tag] = {"events": summary["events"], "now": summary["now"], "draws": summary["draws"],
                    "state": summary["state"], "duration": duration}
        print("frontend", tag, summary["events"], summary["state"] if name != "chatter" else "", flush=True)
    json.dump(out, open(os.path.join(GOLDEN, "frontend_summaries.json"), "w"), indent=1, sort_keys=True)


def malware_golden() -> None:
    """network_simple + malware overlay (config C3): one full trace and two 1 h summaries."""
    out = {}
    for seed, replica, horizon, keep in ((0, 0, 900.0, True), (2, 41, 3600.0, False), (0, 7, 3600.0, False)):
        model = netsimple.MalwareModel(seed, replica)
        model.run(horizon)
        s, arr = summary_of(model, horizon)
        s["infected"] = model.infected.index
        if keep:
            np.savez_compressed(os.path.join(GOLDEN, f"malware_trace_s{seed}_r{replica}_{int(horizon)}s.npz"), trace=arr)
        out[f"{seed}:{replica}:{int(horizon)}"] = s
        model.close()
        print("malware", seed, replica, horizon, s["events"], flush=True)
    json.dump(out, open(os.path.join(GOLDEN, "malware_summaries.json"), "w"), indent=1, sort_keys=True)


def main() -> None:
    os.makedirs(GOLDEN, exist_ok=True)
    if "--machine" in sys.argv:
        machine_golden()
        return
    if "--frontend" in sys.argv:
        frontend_golden()
        return
    if "--malware" in sys.argv:
        malware_golden()
        return
    path = os.path.join(GOLDEN, "netsimple_summaries.json")
    summaries = json.load(open(path)) if os.path.exists(path) else {}
    if "--merge" in sys.argv:
        for name in sorted(os.listdir(GOLDEN)):
            if name.startswith(".part_"):
                summaries.update(json.load(open(os.path.join(GOLDEN, name))))
                os.remove(os.path.join(GOLDEN, name))
        json.dump(summaries, open(path, "w"), indent=1, sort_keys=True)
        return
    # 1. full traces, short horizon
    only_keys = [a for a in sys.argv[1:] if a.count(":") == 2]
    for seed, replica, horizon in (() if only_keys else ((0, 0, 1800.0), (7, 3, 900.0), (0, 0, 600.0), (0, 49999, 600.0))):
        model = netsimple.run_replica(seed, replica, horizon)
        s, arr = summary_of(model, horizon)
        np.savez_compressed(os.path.join(GOLDEN, f"netsimple_trace_s{seed}_r{replica}_{int(horizon)}s.npz"), trace=arr)
        summaries[f"{seed}:{replica}:{int(horizon)}"] = s
        model.close()
        print("trace", seed, replica, horizon, s["events"], flush=True)
    # 2. summaries, 1 h horizon, several streams
    jobs = [(0, r, 3600.0) for r in (0, 1, 2, 3, 63, 49999, 99999)] + [(1, 0, 3600.0), (123456789012, 5, 3600.0)]
    if "--full" in sys.argv:      # config C1 / C2's own horizon: the replicas a whole-job run of C2 is sampled at
        jobs += [(0, r, 43200.0) for r in (0, 1, 2, 3, 63, 99999)]
    only = [a for a in sys.argv[1:] if a.count(":") == 2]     # e.g. 0:63:43200: just these keys (parallel generation)
    for seed, replica, horizon in jobs:
        key = f"{seed}:{replica}:{int(horizon)}"
        if key in summaries or (only and key not in only):
            continue
        model = netsimple.run_replica(seed, replica, horizon, trace_limit=0 if horizon > 10000 else None)
        s, _ = summary_of(model, horizon)
        model.close()
        print("summary", key, s["events"], flush=True)
        if only:        # one process per key: written to its own file, merged by --merge
            json.dump({key: s}, open(os.path.join(GOLDEN, f".part_{key.replace(':', '_')}.json"), "w"))
        else:
            summaries[key] = s
            json.dump(summaries, open(path, "w"), indent=1, sort_keys=True)
    json.dump(summaries, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
