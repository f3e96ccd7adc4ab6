#!/usr/bin/env python
"""
bench.py -- simulated events/sec of batched network_simple replicas on N B200s (BASELINE.json metric).

Workload (config[1] of BASELINE.json): ``network_simple`` (reference examples/network_simple, default 50 endpoints on
a /24) x REPLICAS independent replicas per GPU, replica r of rank g drawing from Philox key (seed, g*REPLICAS + r).
One *step* = ``Simulator.run(SLICE)`` on every replica of the batch (one launch of the persistent engine kernel):
the hot path over one batch.  The batch is first advanced, untimed, by SPINUP simulated seconds so that the timed
steps run in the model's steady state (at t = 0 every workstation is asleep).

What is timed
  value   device time (CUDA events on the engine's stream) of K steps queued back to back, state resident in HBM,
          max over ranks; events = the engine's own count of simulated events (SURVEY.md section 8-d definition).
  e2e     the same K steps driven one by one through the public API from the host: each step launches, waits, and
          reads the telemetry vector back (D2H) -- host wall clock.  Inputs of a step are the scalar duration and
          the kernel arguments (the replica state is resident by design, as in the reference where it stays in the
          interpreter); the model upload / allocation cost of a cold job is reported beside it in config.
  cpu_baseline / --impl reference
          the oracle (the reference's Python model on the greensim shim -- the reference is pure Python and greensim is
          not installable here) run one OS process per host core on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simulated events/sec (batched network_simple replicas)"
# ncu --set full captures committed under profiles/ (6000 replicas x 60 simulated s, one launch):
# (dram__bytes_read.sum + dram__bytes_write.sum) / replicas, per engine kernel
# k_run stages a replica's state once per launch: its traffic goes per launch.  k_run_vote advances the state in place:
# its traffic goes with the simulated time (100000 replicas x 60 simulated s: 97.36 GB read + 16.78 GB written).
NCU_TRAFFIC_BYTES_PER_REPLICA_LAUNCH = {"k_run": (328.9e6 + 160.1e6) / 6000.0}
NCU_TRAFFIC_BYTES_PER_REPLICA_SECOND = {"k_run_vote": (97.36e9 + 16.78e9) / 100000.0 / 60.0}
UNIT = "events/s"
RESULT_OUT = sys.stdout


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle, one process per core
# ---------------------------------------------------------------------------------------------------------------
def _oracle_job(args):
    seed, replica, horizon = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import netsimple as oracle_netsimple
    t0 = time.perf_counter()
    model = oracle_netsimple.run_replica(seed, replica, horizon, trace=False)
    events = model.tracer.num_events
    pops = model.sim.num_pops
    model.close()
    return events, pops, time.perf_counter() - t0


def cpu_baseline(horizon: float, replicas_per_core: int = 1, cores: int = 0, seed: int = 0):
    """Runs whole replicas of the workload on the host cores; returns (events/s aggregate, dict)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    jobs = [(seed, r, horizon) for r in range(cores * replicas_per_core)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        out = pool.map(_oracle_job, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    events = sum(o[0] for o in out)
    pops = sum(o[1] for o in out)
    busy = sum(o[2] for o in out)
    info = {
        "value": events / wall, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{len(jobs)} replicas of network_simple x {horizon:.0f} simulated s from t=0, one process per core "
                  f"(oracle = reference model on the greensim shim, CPython {sys.version_info.major}."
                  f"{sys.version_info.minor})",
        "events": events, "raw_heap_pops_per_s": pops / wall, "events_per_s_per_core": events / busy,
        "wall_s": wall,
    }
    return events / wall, info


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            try:
                sm.append(float(row[1]))
                smax.append(float(row[2]))
            except (ValueError, IndexError):
                continue
            for name, cell in zip(names, row[5:9]):
                if cell.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = []
    info = None
    for _ in range(args.warmup):
        cpu_baseline(args.cpu_horizon / 4, 1, cores)
    t_events = 0
    t_wall = 0.0
    for _ in range(args.steps):
        _, info = cpu_baseline(args.cpu_horizon, 1, cores)
        per_step.append(info["wall_s"])
        t_events += info["events"]
        t_wall += info["wall_s"]
    value = t_events / t_wall
    info["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_wall / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "network_simple (50 endpoints, /24), whole replicas on host cores, one process per core",
                   "sample_horizon_s": args.cpu_horizon, "note": "reference is pure Python on greensim; greensim is "
                   "not installable offline, so the reference model runs on the oracle's greensim-compatible shim"},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


def replica_histograms(engine, args, L, netsimple, comm, world):
    """Config 3's output: histograms over ALL replicas of the job of beacons per replica, bytes that left the local
    network and the time of the first beacon (`itsb_histogram`: one privatised kernel per statistic, bins summed over
    the ranks with ncclAllReduce when there is a communicator).  Returns (dict, milliseconds for the three)."""
    if args.workload not in ("malware", "backdoor"):
        return None, None
    horizon_us = int((args.spinup + (2 * args.steps + args.warmup) * args.slice) * 1e6)
    spec = {"beacons_per_replica": (L.tm_user(netsimple.STAT_BEACONS), 0, 8, 64, False),
            "exfiltrated_bytes_per_replica": (L.tm_user(netsimple.STAT_EXFIL_BYTES), 0, 800, 64, False),
            "first_beacon_us": (L.tm_user(netsimple.STAT_FIRST_BEACON_US), 0, max(1, horizon_us // 64), 64, True)}
    out = {}
    t0 = time.perf_counter()
    for name, (slot, lo, width, nbins, nonzero) in spec.items():
        bins = engine.histogram(slot, lo, width, nbins, only_nonzero=nonzero, comm=comm)
        out[name] = {"lo": lo, "bin_width": width, "bins": [int(b) for b in bins]}
    ms = 1e3 * (time.perf_counter() - t0)
    assert sum(out["beacons_per_replica"]["bins"]) == args.replicas * world, "histogram does not cover every replica"
    return out, ms


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=100000, help="replicas per GPU")
    ap.add_argument("--slice", type=float, default=120.0, help="simulated seconds per step")
    ap.add_argument("--spinup", type=float, default=1800.0, help="untimed simulated seconds before warm-up")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-horizon", type=float, default=1800.0, help="simulated seconds per CPU-baseline replica")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the untimed full-size self-check after the timed steps")
    ap.add_argument("--arena-nodes", type=int, default=12288)
    ap.add_argument("--max-events", type=int, default=0, help="future-event pool slots per replica (0 = model default)")
    ap.add_argument("--max-procs", type=int, default=0)
    ap.add_argument("--max-sockets", type=int, default=0)
    ap.add_argument("--net-events", type=int, default=0,
                    help="per-replica capacity of the network_event log in HBM (0 = off): the logging-on variant")
    ap.add_argument("--connections", type=int, default=0,
                    help="per-replica capacity of the connection-record log (Socket.log_pair) in HBM (0 = off); with "
                         "--net-events this is BASELINE.json config 5, full logging")
    ap.add_argument("--connection-window", type=int, default=8)
    ap.add_argument("--workload", default="network_simple", choices=["network_simple", "malware", "backdoor", "large_lan"],
                    help="malware = network_simple + the beaconing-malware overlay (BASELINE.json config 3); "
                         "backdoor = the overlay with the parameters of examples/backdoor/backdoor.py; "
                         "large_lan = 1024 endpoints + 16 forwarding routers per replica (config 4; use "
                         "--replicas 10000 --slice 2 --spinup 6)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner) goes to stderr
    global RESULT_OUT
    RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import numpy as np
    from itsim_b200 import netsimple, _lib as L
    from itsim_b200.engine import summarize

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- cold job through the public API: compile, upload, allocate, initialise --------------------------------
    t_cold0 = time.perf_counter()
    if args.workload == "large_lan":
        from itsim_b200 import machine_models
        sim = machine_models.large_lan()
    else:
        build = {"network_simple": netsimple.build, "malware": netsimple.build_malware,
                 "backdoor": netsimple.build_backdoor}[args.workload]
        sim = build(connection_records=args.connections, connection_window=args.connection_window)
        sim.caps["arena_nodes"] = args.arena_nodes
    sim.caps["net_event_records"] = args.net_events
    for key, val in (("max_events", args.max_events), ("max_procs", args.max_procs), ("max_sockets", args.max_sockets)):
        if val:
            sim.caps[key] = val
    engine = sim.run_replicas(args.replicas, seed=args.seed, first_replica_id=rank * args.replicas, device=local_rank)
    t_cold = time.perf_counter() - t_cold0
    state = engine.state_bytes()
    kernel = engine.kernel_name()
    h2d_model = sum(b.nbytes for b in engine.compiled.blobs())

    spin_events = engine.run(args.spinup) if args.spinup > 0 else 0
    spin_ms = engine.last_run_ms()
    warm_events = 0
    for _ in range(args.warmup):
        warm_events += engine.run(args.slice)
    log_records_before = int(engine.telemetry()[L.TM_USER0 + 7])
    conn_records_before = int(engine.telemetry()[L.TM_USER0 + 6]) if args.connections else 0

    # ---- timed region 1: K steps queued back to back, device time ------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = engine.launch_count()
    barrier()
    for _ in range(args.steps):
        engine.run_async(args.slice)
    events_dev = engine.sync()
    ms_dev = engine.last_run_ms()
    barrier()
    launches_dev = engine.launch_count() - launches0
    ms_dev_max = allmax(ms_dev)
    events_all = allsum(float(events_dev))
    value = events_all / (ms_dev_max * 1e-3)

    # ---- timed region 2: the same K steps, one by one from the host, telemetry read back every step ------------
    barrier()
    t0 = time.perf_counter()
    events_e2e = 0
    for _ in range(args.steps):
        events_e2e += engine.run(args.slice)
        tm = engine.telemetry()
    wall_e2e = time.perf_counter() - t0
    wall_job = time.perf_counter() - t_cold0        # everything since the model was built: compile, upload, allocation,
    events_job = spin_events + warm_events + events_dev + events_e2e      # spin-up, warm-up, both timed regions
    barrier()
    clocks = sampler.stop()
    wall_e2e_max = allmax(wall_e2e)
    e2e_value = allsum(float(events_e2e)) / wall_e2e_max

    # ---- the one collective on the path: telemetry histogram all-reduce ----------------------------------------------
    allreduce_ms = None
    if dist is not None:
        import ctypes as C
        uid = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            rc = engine._lib.itsb_nccl_unique_id(uid.ctypes.data, 128)
            assert rc == 0, rc
        box = [uid.tobytes()]
        dist.broadcast_object_list(box, src=0)
        uid = np.frombuffer(box[0], dtype=np.uint8).copy()
        comm = C.c_void_p()
        rc = engine._lib.itsb_nccl_comm_init(engine._h, world, rank, uid.ctypes.data, C.byref(comm))
        assert rc == 0, engine._lib.itsb_last_error(engine._h)
        local = engine.telemetry()
        engine.allreduce_telemetry(comm)          # warm-up (communicator setup)
        t0 = time.perf_counter()
        total = engine.allreduce_telemetry(comm)
        allreduce_ms = 1e3 * (time.perf_counter() - t0)
        gathered = [None] * world
        dist.all_gather_object(gathered, local.tolist())
        assert np.array_equal(total, np.sum(np.array(gathered, dtype=np.uint64), axis=0)), "allreduce mismatch"
        hists, hist_ms = replica_histograms(engine, args, L, netsimple, comm, world)
        engine._lib.itsb_nccl_comm_destroy(comm)
    else:
        total = engine.telemetry()
        hists, hist_ms = replica_histograms(engine, args, L, netsimple, None, 1)

    # ---- full-size self-check (untimed): conservation laws on the job's counters, and sampled replicas of the batch
    # ---- equal the same replica ids advanced alone through the same sequence of run() calls -----------------------
    checks = None
    if rank == 0 and not args.no_verify:
        tm_job = np.asarray(total, dtype=np.uint64)
        k0 = L.TM_KIND0
        checks = {"events_equal_sum_of_kinds": int(tm_job[:9].sum()) == int(summarize(tm_job)["events"])}
        if args.workload == "large_lan":      # a forwarding router's transit packets are neither delivered nor dropped
            checks["delivered_plus_dropped_at_most_deliver_events"] = \
                int(tm_job[L.TM_DELIVERED]) + int(tm_job[L.TM_DROPPED]) <= int(tm_job[k0 + L.EV_DELIVER])
        else:
            checks["delivered_plus_dropped_equal_deliver_events"] = \
                int(tm_job[L.TM_DELIVERED]) + int(tm_job[L.TM_DROPPED]) == int(tm_job[k0 + L.EV_DELIVER])
            checks["tx_begin_equal_packets_sent"] = int(tm_job[k0 + L.EV_TX_BEGIN]) == int(tm_job[L.TM_PACKETS_SENT])
            checks["latency_histogram_covers_every_transmission"] = \
                int(tm_job[L.TM_LAT0:L.TM_LAT0 + L.LAT_BINS].sum()) == int(tm_job[k0 + L.EV_TX_BEGIN])
        sample = sorted({0, args.replicas // 2, args.replicas - 1})
        same = []
        for r in sample:
            alone = sim.run_replicas(1, seed=args.seed, first_replica_id=rank * args.replicas + r, device=local_rank)
            if args.spinup > 0:
                alone.run(args.spinup)
            for _ in range(args.warmup + 2 * args.steps):
                alone.run(args.slice)
            same.append(bool(np.array_equal(alone.replica_telemetry(0), engine.replica_telemetry(r))))
            alone.close()
        checks["replicas_equal_themselves_run_alone"] = dict(zip(map(str, sample), same))
        checks["all"] = all(v if isinstance(v, bool) else all(v.values()) for v in checks.values())
        assert checks["all"], checks

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        events_per_step_gpu = events_dev / args.steps
        log_records = (int(total[L.TM_USER0 + 7]) // max(1, world) - log_records_before) / (2.0 * args.steps)
        # S_in + S_out per launch, plus B_log = 32 B per network_event record written in a step (0 when the log is off)
        conn_records = ((int(total[L.TM_USER0 + 6]) // max(1, world) - conn_records_before) / (2.0 * args.steps)
                        if args.connections else 0.0)
        algo_bytes = 2.0 * state["hot"] * args.replicas + 32.0 * log_records + 40.0 * conn_records
        launch_s = ms_dev * 1e-3 / args.steps
        achieved = algo_bytes / launch_s / 1e9
        traffic = None
        if kernel in NCU_TRAFFIC_BYTES_PER_REPLICA_LAUNCH:
            traffic = NCU_TRAFFIC_BYTES_PER_REPLICA_LAUNCH[kernel] * args.replicas
        elif kernel in NCU_TRAFFIC_BYTES_PER_REPLICA_SECOND:
            traffic = NCU_TRAFFIC_BYTES_PER_REPLICA_SECOND[kernel] * args.replicas * args.slice
        cpu = None
        if not args.no_cpu_baseline:
            _, cpu = cpu_baseline(args.cpu_horizon, 1, 0, args.seed)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{args.workload} "
                            f"({'1024 endpoints + 16 routers' if args.workload == 'large_lan' else '50 endpoints, /24'})"
                            f" x {args.replicas} replicas per GPU, "
                            f"step = run({args.slice:g} simulated s) on every replica after {args.spinup:g} s spin-up",
                "replicas_per_gpu": args.replicas, "slice_s": args.slice, "spinup_s": args.spinup, "seed": args.seed,
                "events_per_step_per_gpu": events_per_step_gpu,
                "l2": "inputs larger than L2 (replica state %.2f GB per GPU vs 126 MB L2)" %
                      (args.replicas * (state["hot"] + state["arena"]) / 1e9),
                "hot_bytes_per_replica": state["hot"], "arena_bytes_per_replica": state["arena"],
                "net_event_records_per_step_per_gpu": log_records,
                "connection_records_per_step_per_gpu": conn_records,
                "cold_job_s": t_cold, "cold_job_h2d_bytes": h2d_model,
                "spinup_events": spin_events, "spinup_ms": spin_ms,
                "telemetry_allreduce_ms": allreduce_ms,
                "telemetry": {k: v for k, v in summarize(total).items() if k != "latency_hist"},
                "replica_histograms": hists, "replica_histograms_ms": hist_ms,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 320, "d2h_bytes_per_step": 8 * L.TM_SIZE + 24,
                    "note": "per step: kernel arguments + ticket reset H2D; telemetry vector, event count and status D2H",
                    "whole_job": {"value": events_job / wall_job, "unit": UNIT, "wall_s": wall_job, "events": events_job,
                                  "note": "rank 0, host clock from model construction to the last telemetry read: "
                                          "compile + table upload + replica allocation / initialisation + spin-up + "
                                          "warm-up + both timed regions, clock sampler start included"}},
            "gpu_launches": launches_dev,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "dram_gbs": (traffic / launch_s / 1e9) if traffic else None,
                "peak_source": peak_src, "kernel": kernel,
                "traffic_note": "dram__bytes_read+write per launch, scaled from this kernel's ncu --set full capture under "
                                "profiles/ (k_run: 6000 replicas x 60 s, per replica and launch; k_run_vote: 100000 "
                                "replicas x 60 s, per replica and simulated second); dram_gbs = traffic / launch time",
                "algorithmic_bytes_per_launch": algo_bytes,
                "note": "algorithmic bytes = per-replica state read and written once per run (S_in+S_out, SURVEY 8-d); "
                        "neither engine kernel moves bytes at the HBM rate: k_run is bound by instruction supply (GPC-level "
                        "instruction-cache requests at 93% of peak), k_run_vote by memory latency (85% of stall samples "
                        "wait on loads, DRAM at ~13% of peak); profiles/README.md",
            },
            "clocks": clocks,
            "checks": checks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    engine.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
