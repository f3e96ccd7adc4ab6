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
          not installable here) run one OS process per host core on a bounded sample of the same workload; taken BEFORE
          any GPU work of the process, so that the engine's host threads do not share the cores with it.

Multi-GPU: no PyTorch anywhere.  Under a launcher (``python -m torch.distributed.run`` exports RANK / LOCAL_RANK /
WORLD_SIZE) every rank drives its GPU through the C ABI and the ranks meet in NCCL (itsim_b200.distributed.RankEnv:
the NCCL id travels through a file, barrier / max / sum are ncclAllReduce of a few uint64).  ``--gpus N`` without a
launcher runs the same job from ONE process, one host thread per GPU (itsim_b200.distributed.ShardedEngine).
``--scaling strong --total-replicas T`` splits a fixed job of T replicas over the GPUs instead of giving each
``--replicas``.
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
# ncu --set full captures committed under profiles/ (r2_*_ncu_full.txt): dram__bytes_read.sum + dram__bytes_write.sum
# of one launch.  k_run stages a replica's state once per launch: its traffic goes per launch (6000 replicas x 60 s).  The
# lane-per-replica kernels advance the state in place: their traffic goes with the simulated time (100000 replicas x 60
# simulated s per launch: k_run_pool 72.37 GB read + 23.99 GB written, k_run_vote 61.00 + 22.15 GB).
NCU_TRAFFIC_BYTES_PER_REPLICA_LAUNCH = {"k_run": (328.9e6 + 160.1e6) / 6000.0}
NCU_TRAFFIC_BYTES_PER_REPLICA_SECOND = {"k_run_pool": (72.37e9 + 23.99e9) / 100000.0 / 60.0,
                                        "k_run_vote": (61.00e9 + 22.15e9) / 100000.0 / 60.0}
NCU_TRAFFIC_NOTE = ("dram__bytes_read+write per launch, scaled from this kernel's ncu --set full capture under profiles/ "
                    "(k_run: 6000 replicas x 60 s, per replica and launch; k_run_pool / k_run_vote: 100000 replicas x 60 s, "
                    "per replica and simulated second); dram_gbs = traffic / launch time")
ROOFLINE_NOTE = ("algorithmic bytes = per-replica state read and written once per run (S_in+S_out, SURVEY 8-d) plus the "
                 "records a step logs; the engine kernels do not move bytes at the HBM rate: every one of them runs at the "
                 "rate its SMs' instruction-cache misses are served by the GPC-level cache (gcc requests at 91-93% of peak "
                 "in every capture), k_run_pool with 70% of its stall samples waiting on dependent loads of per-replica "
                 "state on top; profiles/README.md")
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


def shim_switch_cost_us(n: int = 20000) -> float:
    """Cost of one process switch of the greensim shim (the oracle's coroutines are OS threads handing a baton to each
    other -- greenlet is not installable here -- and that hand-off, not the model, is most of the CPU arm's time):
    microseconds per heap pop of two processes that do nothing but ``advance(1.0)``."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import greensim

    def idle():
        while True:
            greensim.advance(1.0)

    sim = greensim.Simulator()
    sim.add(idle)
    sim.add(idle)
    sim.run(50.0)
    t0 = time.perf_counter()
    before = sim.num_pops
    sim.run(float(n // 2))
    dt = time.perf_counter() - t0
    pops = sim.num_pops - before
    sim._clear()
    return 1e6 * dt / max(1, pops)


def cpu_baseline(horizon: float, replicas_per_core: int = 1, cores: int = 0, seed: int = 0, pool=None):
    """Runs whole replicas of the workload on the host cores; returns (events/s aggregate, dict).  The worker processes
    are started and have imported the model before the clock starts (an untimed 1-second replica each): interpreter
    start-up and imports were 10-45 % of the short samples of round 1, cold or warm page cache deciding which."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    jobs = [(seed, r, horizon) for r in range(cores * replicas_per_core)]
    own = pool is None
    t_start = time.perf_counter()
    if own:
        pool = mp.get_context("spawn").Pool(cores)
        pool.map(_oracle_job, [(seed, r, 1.0) for r in range(cores)], chunksize=1)
    startup = time.perf_counter() - t_start
    try:
        t0 = time.perf_counter()
        out = pool.map(_oracle_job, jobs, chunksize=1)
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    events = sum(o[0] for o in out)
    pops = sum(o[1] for o in out)
    busy = sum(o[2] for o in out)
    info = {
        "value": events / wall, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{len(jobs)} replicas of network_simple x {horizon:.0f} simulated s from t=0, one process per core "
                  f"(oracle = reference model on the greensim shim, CPython {sys.version_info.major}."
                  f"{sys.version_info.minor})",
        "events": events, "raw_heap_pops_per_s": pops / wall, "events_per_s_per_core": events / busy,
        "heap_pops_per_event": pops / max(1, events), "shim_switch_cost_us": shim_switch_cost_us(),
        "wall_s": wall, "worker_startup_s_untimed": startup,
        "why_not_the_12_h_horizon": "config C1's horizon (43 200 s) is ~3.8e6 events = several minutes per replica on "
                                    "this shim; the contract bounds the CPU sample to 10-30 s, so a replica is run for "
                                    "the first sample_horizon seconds (the workstations wake up during the first ~600 s;"
                                    " events/s is events over wall time, so the quiet start costs little of either)",
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
    horizon = args.cpu_horizon / 2          # per step: sized so that --steps 20 --warmup 5 ends within a few minutes
    import multiprocessing as mp
    t_events = 0
    t_wall = 0.0
    with mp.get_context("spawn").Pool(cores) as pool:      # one set of workers for the whole arm
        for _ in range(args.warmup):
            cpu_baseline(horizon / 4, 1, cores, pool=pool)
        for _ in range(args.steps):
            _, info = cpu_baseline(horizon, 1, cores, pool=pool)
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
                   "sample_horizon_s": horizon, "note": "reference is pure Python on greensim; greensim is "
                   "not installable offline, so the reference model runs on the oracle's greensim-compatible shim"},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=RESULT_OUT, flush=True)


class _Job:
    """The replicas this PROCESS advances -- one engine, or one engine per GPU behind ShardedEngine (``--gpus N`` without
    a launcher) -- and the reductions over the ranks of the job (NCCL through RankEnv; nothing with one process)."""

    def __init__(self, sim, args, env, threads: int):
        from itsim_b200.distributed import ShardedEngine
        from itsim_b200.sharding import shard
        self.env, self.threads = env, threads
        world = threads if threads > 1 else env.world
        rank = 0 if threads > 1 else env.rank
        if args.scaling == "strong":
            self.total = args.total_replicas or args.replicas
            self.first, self.count = shard(self.total, rank, world)
        else:
            self.total = args.replicas * world
            self.first, self.count = rank * args.replicas, args.replicas
        if threads > 1:
            self.first, self.count = 0, self.total
            self.sharded = ShardedEngine(sim, self.total, seed=args.seed, devices=list(range(threads)))
            self.engine = self.sharded.engines[0]
            self.local_count = self.sharded.shards[0][1]
        else:
            self.sharded = None
            self.engine = sim.run_replicas(self.count, seed=args.seed, first_replica_id=self.first, device=env.local_rank)
            self.local_count = self.count
            env.join(self.engine)
        self.world = world
        self.logging = bool(args.connections or args.net_events)
        self.records = {"net": 0, "conn": 0, "net_dropped": 0, "conn_dropped": 0, "d2h_bytes": 0, "drains": 0,
                        "pack_ms": 0.0, "pack_ms_last": 0.0}

    # -- the engines of this process --------------------------------------------------------------------------------
    def _engines(self):
        return self.sharded.engines if self.sharded else [self.engine]

    def run(self, dt):
        return self.sharded.run(dt) if self.sharded else self.engine.run(dt)

    def run_async(self, dt):
        (self.sharded or self.engine).run_async(dt)

    def sync(self):
        return (self.sharded or self.engine).sync()

    def last_run_ms(self):
        return (self.sharded or self.engine).last_run_ms()

    def launch_count(self):
        return (self.sharded or self.engine).launch_count()

    def drain_begin(self):
        for e in self._engines():
            e.drain_begin()

    def drain_end(self):
        """Waits for the oldest pending drain of every engine; what came home is counted (and touched: one pass over
        the timestamps, so that the pages really are in host memory)."""
        self.records["pack_ms_last"] = 0.0
        for e in self._engines():
            d = e.drain_end()
            self.records["pack_ms"] += d["pack_ms"]
            self.records["pack_ms_last"] = max(self.records["pack_ms_last"], d["pack_ms"])
            self.records["net"] += len(d["net"]); self.records["conn"] += len(d["conn"])
            self.records["net_dropped"] += d["net_dropped"]; self.records["conn_dropped"] += d["conn_dropped"]
            self.records["d2h_bytes"] += d["d2h_bytes"]; self.records["drains"] += 1
            if len(d["conn"]):
                assert float(d["conn"]["t_end"].max()) >= 0.0
        return self.records

    def advance(self, total, chunk):
        """Untimed: `total` simulated seconds in `chunk`-second runs, the record logs emptied after each one when the
        model keeps them (they are windows sized for a step, not for the whole horizon)."""
        events, left = 0, total
        while left > 1e-9:
            dt = min(chunk, left) if self.logging else left
            events += self.run(dt)
            if self.logging:
                self.drain_begin()
                self.drain_end()
            left -= dt
        return events

    # -- over the ranks ------------------------------------------------------------------------------------------------
    def barrier(self):
        if not self.sharded:
            self.env.barrier()

    def all_max(self, x):
        return x if self.sharded else self.env.max_float(x)

    def all_sum(self, x):
        return int(x) if self.sharded else self.env.sum_int(int(x))

    def telemetry_job(self):
        if self.sharded:
            return self.sharded.telemetry()
        return self.engine.allreduce_telemetry(self.env.comm) if self.env.world > 1 else self.engine.telemetry()

    def histogram(self, slot, lo, width, nbins, only_nonzero):
        if self.sharded:
            return self.sharded.histogram(slot, lo, width, nbins, only_nonzero)
        return self.engine.histogram(slot, lo, width, nbins, only_nonzero=only_nonzero,
                                     comm=self.env.comm if self.env.world > 1 else None)

    def close(self):
        if self.sharded:
            self.sharded.close()
        else:
            self.env.close()
            self.engine.close()


def replica_histograms(job, args, L, netsimple):
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
        bins = job.histogram(slot, lo, width, nbins, nonzero)
        out[name] = {"lo": lo, "bin_width": width, "bins": [int(b) for b in bins]}
    ms = 1e3 * (time.perf_counter() - t0)
    assert sum(out["beacons_per_replica"]["bins"]) == job.total, "histogram does not cover every replica"
    return out, ms


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicas", type=int, default=100000, help="replicas per GPU (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --replicas on every GPU; strong: --total-replicas split over the GPUs")
    ap.add_argument("--total-replicas", type=int, default=0, help="size of the job under --scaling strong")
    ap.add_argument("--slice", type=float, default=120.0, help="simulated seconds per step")
    ap.add_argument("--spinup", type=float, default=1800.0, help="untimed simulated seconds before warm-up")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-horizon", type=float, default=3600.0, help="simulated seconds per CPU-baseline replica")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the untimed full-size self-check after the timed steps")
    ap.add_argument("--arena-nodes", type=int, default=12288)
    ap.add_argument("--max-events", type=int, default=0, help="future-event heap slots per replica (0 = model default)")
    ap.add_argument("--max-procs", type=int, default=0)
    ap.add_argument("--max-sockets", type=int, default=0)
    ap.add_argument("--net-events", type=int, default=0,
                    help="per-replica capacity of the network_event log in HBM (0 = off): the logging-on variant")
    ap.add_argument("--connections", type=int, default=0,
                    help="per-replica capacity of the connection-record log (Socket.log_pair) in HBM (0 = off); with "
                         "--net-events this is BASELINE.json config 5, full logging.  Both logs are drained to the host "
                         "after every step (itsb_drain_begin / itsb_drain_end), inside the timed regions")
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

    import numpy as np
    from itsim_b200 import netsimple, _lib as L
    from itsim_b200.distributed import RankEnv
    from itsim_b200.engine import summarize

    env = RankEnv()
    rank, world = env.rank, env.world
    threads = args.gpus if (world == 1 and args.gpus > 1) else 1      # no launcher: one process, one thread per GPU
    n_gpus = threads if threads > 1 else world

    # ---- the CPU arm first, before this process starts any GPU work (its threads would share the cores) -------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        _, cpu = cpu_baseline(args.cpu_horizon, 1, 0, args.seed)

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
    job = _Job(sim, args, env, threads)
    engine = job.engine
    t_cold = time.perf_counter() - t_cold0
    state = engine.state_bytes()
    kernel = engine.kernel_name()
    h2d_model = sum(b.nbytes for b in engine.compiled.blobs())

    spin_events = job.advance(args.spinup, args.slice) if args.spinup > 0 else 0
    spin_ms = job.last_run_ms()
    warm_events = 0
    for _ in range(args.warmup):
        warm_events += job.advance(args.slice, args.slice)
    records_before = dict(job.records)

    # ---- timed region 1: K steps, device time ------------------------------------------------------------------------
    # Without record logs the K launches are queued back to back and timed by one pair of CUDA events on the engine's
    # stream.  With them, every step is followed by its drain (the logs are windows sized for a step), so the steps are
    # timed one by one -- the engine kernel and the drain's count / scan / gather kernels each by their own CUDA events
    # -- and the device time of the region is their sum.  The copy home of the records belongs to e2e (region 2).
    sampler = ClockSampler(env.local_rank)
    sampler.start()
    launches0 = job.launch_count()
    job.barrier()
    if not job.logging:
        for _ in range(args.steps):
            job.run_async(args.slice)
        events_dev = job.sync()
        ms_dev = job.last_run_ms()
    else:
        events_dev, ms_dev = 0, 0.0
        for _ in range(args.steps):
            events_dev += job.run(args.slice)
            ms_dev += job.last_run_ms()
            job.drain_begin()
            ms_dev += job.drain_end()["pack_ms_last"]
    job.barrier()
    launches_dev = job.launch_count() - launches0
    ms_dev_max = job.all_max(ms_dev)
    events_all = job.all_sum(events_dev)
    value = events_all / (ms_dev_max * 1e-3)

    # ---- timed region 2: the same K steps, one by one from the host through the public API: every step launches,
    # ---- waits, reads the telemetry vector back, and -- when the model keeps record logs -- brings every record of the
    # ---- step home (the copy of step k overlaps the kernel of step k + 1) ------------------------------------------
    job.barrier()
    records_mid = dict(job.records)
    t0 = time.perf_counter()
    events_e2e = 0
    pending = False
    for _ in range(args.steps):
        job.run_async(args.slice)
        if pending:
            job.drain_end()
        events_e2e += job.sync()
        tm = engine.telemetry()
        if job.logging:
            job.drain_begin()
            pending = True
    if pending:
        job.drain_end()
    wall_e2e = time.perf_counter() - t0
    wall_job = time.perf_counter() - t_cold0        # everything since the model was built: compile, upload, allocation,
    events_job = spin_events + warm_events + events_dev + events_e2e      # spin-up, warm-up, both timed regions
    job.barrier()
    clocks = sampler.stop()
    wall_e2e_max = job.all_max(wall_e2e)
    e2e_value = job.all_sum(events_e2e) / wall_e2e_max
    d2h_records_per_step = (job.records["d2h_bytes"] - records_mid["d2h_bytes"]) / max(1, args.steps)

    # ---- the one collective on the path: telemetry histogram all-reduce ----------------------------------------------
    allreduce_ms = None
    if n_gpus > 1:
        job.telemetry_job()                       # warm-up (first collective on the communicator)
        t0 = time.perf_counter()
        total = job.telemetry_job()
        allreduce_ms = 1e3 * (time.perf_counter() - t0)
        if not job.sharded:                       # the reduced vector is the sum of the ranks' own vectors, slot by slot
            local = engine.telemetry()
            for slot in (0, L.TM_KIND0 + L.EV_DELIVER, L.TM_BYTES_SENT, L.TM_PACKETS_RECEIVED):
                assert int(total[slot]) == job.all_sum(int(local[slot])), "allreduce mismatch"
    else:
        total = job.telemetry_job()
    hists, hist_ms = replica_histograms(job, args, L, netsimple)

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
        if job.logging:
            checks["no_record_dropped"] = job.records["net_dropped"] == 0 and job.records["conn_dropped"] == 0
            if n_gpus == 1:
                checks["every_record_produced_came_home"] = \
                    int(tm_job[L.TM_USER0 + 6]) == job.records["conn"] and int(tm_job[L.TM_USER0 + 7]) == job.records["net"]
        sample = sorted({0, job.local_count // 2, job.local_count - 1})
        same = []
        for r in sample:
            alone = sim.run_replicas(1, seed=args.seed, first_replica_id=job.first + r, device=env.local_rank)
            alone_job_logging = job.logging
            calls = ([args.spinup] if args.spinup > 0 else []) + [args.slice] * (args.warmup + 2 * args.steps)
            for dt in calls:
                left = dt
                while left > 1e-9:              # the same sequence of run() calls the batch went through
                    step = min(args.slice, left) if alone_job_logging else left
                    alone.run(step)
                    if alone_job_logging:
                        alone.drain()
                    left -= step
            same.append(bool(np.array_equal(alone.replica_telemetry(0), engine.replica_telemetry(r))))
            alone.close()
        checks["replicas_equal_themselves_run_alone"] = dict(zip(map(str, sample), same))
        checks["all"] = all(v if isinstance(v, bool) else all(v.values()) for v in checks.values())
        assert checks["all"], checks

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        per_gpu = job.local_count
        events_per_step_gpu = events_dev / args.steps / (threads if threads > 1 else 1)
        steps_logged = 2.0 * args.steps
        log_records = (job.records["net"] - records_before["net"]) / steps_logged / (threads if threads > 1 else 1)
        conn_records = (job.records["conn"] - records_before["conn"]) / steps_logged / (threads if threads > 1 else 1)
        # S_in + S_out per launch, plus B_log = 32 B per network_event record and 40 B per connection record of a step
        algo_bytes = 2.0 * state["hot"] * per_gpu + 32.0 * log_records + 40.0 * conn_records
        launch_s = ms_dev * 1e-3 / args.steps
        achieved = algo_bytes / launch_s / 1e9
        traffic = None
        if kernel in NCU_TRAFFIC_BYTES_PER_REPLICA_LAUNCH:
            traffic = NCU_TRAFFIC_BYTES_PER_REPLICA_LAUNCH[kernel] * per_gpu
        elif kernel in NCU_TRAFFIC_BYTES_PER_REPLICA_SECOND:
            traffic = NCU_TRAFFIC_BYTES_PER_REPLICA_SECOND[kernel] * per_gpu * args.slice
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev_max / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"{args.workload} "
                            f"({'1024 endpoints + 16 routers' if args.workload == 'large_lan' else '50 endpoints, /24'})"
                            + (f" x {args.replicas} replicas per GPU, " if args.scaling == "weak" else
                               f" x {job.total} replicas in all, split over {n_gpus} GPU(s) ({per_gpu} on rank 0), ")
                            + f"step = run({args.slice:g} simulated s) on every replica after {args.spinup:g} s spin-up",
                "replicas_per_gpu": per_gpu, "replicas_total": job.total, "slice_s": args.slice,
                "spinup_s": args.spinup, "seed": args.seed,
                "processes": "one per GPU (launcher), NCCL between them" if world > 1 else
                             ("one, a host thread per GPU" if threads > 1 else "one"),
                "events_per_step_per_gpu": events_per_step_gpu,
                "l2": "inputs larger than L2 (replica state %.2f GB per GPU vs 126 MB L2)" %
                      (per_gpu * (state["hot"] + state["arena"]) / 1e9),
                "hot_bytes_per_replica": state["hot"], "arena_bytes_per_replica": state["arena"],
                "net_event_records_per_step_per_gpu": log_records,
                "connection_records_per_step_per_gpu": conn_records,
                "records": dict(job.records) if job.logging else None,
                "cold_job_s": t_cold, "cold_job_h2d_bytes": h2d_model,
                "spinup_events": spin_events, "spinup_ms": spin_ms,
                "telemetry_allreduce_ms": allreduce_ms,
                "telemetry": {k: v for k, v in summarize(total).items() if k != "latency_hist"},
                "replica_histograms": hists, "replica_histograms_ms": hist_ms,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 320,
                    "d2h_bytes_per_step": 8 * L.TM_SIZE + 24 + d2h_records_per_step,
                    "note": "per step: kernel arguments + ticket reset H2D; telemetry vector, event count and status D2H"
                            + ("; plus every network_event / connection record of the step, packed on the device and "
                               "copied to pinned host memory (itsb_drain_*), the copy overlapping the next step's kernel"
                               if job.logging else ""),
                    "whole_job": {"value": events_job / wall_job, "unit": UNIT, "wall_s": wall_job, "events": events_job,
                                  "note": "rank 0, host clock from model construction to the last telemetry read: "
                                          "compile + table upload + replica allocation / initialisation + spin-up + "
                                          "warm-up + both timed regions, clock sampler start included"}},
            "gpu_launches": launches_dev,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "dram_gbs": (traffic / launch_s / 1e9) if traffic else None,
                "traffic_over_algorithmic": (traffic / algo_bytes) if traffic else None,
                "peak_source": peak_src, "kernel": kernel,
                "traffic_note": NCU_TRAFFIC_NOTE,
                "algorithmic_bytes_per_launch": algo_bytes,
                "note": ROOFLINE_NOTE,
            },
            "clocks": clocks,
            "checks": checks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), file=RESULT_OUT, flush=True)
    job.close()


if __name__ == "__main__":
    main()
