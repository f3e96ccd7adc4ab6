"""
Multi-GPU launcher: the replica set of one model split over the GPUs of a box.

Replicas share nothing: shard ``g`` of ``G`` owns a contiguous range of global replica ids
(:func:`itsim_b200.sharding.shard`), the Philox key carries the global id, so a replica's trajectory does not depend on
the split, and there is no exchange while simulating.  The one collective is the final sum of the telemetry vectors /
histogram bins: ``ncclAllReduce`` over NVLink inside ``itsb_allreduce_telemetry`` / ``itsb_histogram`` /
``itsb_allreduce_u64`` (NCCL is resolved by the library itself; no PyTorch anywhere on this path).

Two ways to place the shards, same results:

* :class:`ShardedEngine` -- ONE process, one host thread per GPU.  ctypes releases the GIL inside every ``itsb_*`` call
  and engines on different GPUs are independent (include/itsb.h "Threading"), so the threads only exist to issue the
  launches concurrently and to enter the NCCL calls together.  This is what a user script calls:
  ``run_sharded(sim, n_total, seed, until)``.
* :class:`RankEnv` -- one process per GPU under an external launcher (``python -m torch.distributed.run`` sets
  ``RANK`` / ``LOCAL_RANK`` / ``WORLD_SIZE`` / ``MASTER_PORT``; nothing of torch is imported): the environment is
  read, the 128-byte NCCL unique id travels from rank 0 to the others through a file in the box's temp directory (all
  ranks are on one node), and barrier / max / sum over the ranks are NCCL all-reduces of a few uint64.

Reference counterpart: none -- the reference is single-threaded; a user would start N
interpreters with N seeds.
"""
import ctypes as C
import os
import struct
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, List, Optional, Sequence

import numpy as np

from . import _lib as L
from .sharding import shard

NCCL_SUM, NCCL_MAX = 0, 2          # ncclRedOp_t (nccl.h)


class ShardedEngine:
    """``n_total`` replicas of ``sim`` over ``devices`` (default: every visible GPU), one host thread per device.

    ``reduce`` says how the shards' telemetry is summed: ``"nccl"`` (default with more than one device: one
    communicator over the devices, ``ncclAllReduce`` on each engine's stream) or ``"host"`` (vectors read back and added
    on the host -- what the CPU test-suite uses with the host emulation, which has no NCCL)."""

    def __init__(self, sim: Any, n_total: int, seed: int = 0, devices: Optional[Sequence[int]] = None,
                 first_replica_id: int = 0, reduce: Optional[str] = None, lib: Optional[C.CDLL] = None) -> None:
        from .engine import Engine
        self._lib = lib if lib is not None else (sim._lib if getattr(sim, "_lib", None) is not None else L.load())
        if devices is None:
            devices = list(range(device_count()))
        if not devices:
            raise L.EngineError(-1, "ShardedEngine", "no CUDA device")
        self.devices = list(devices)
        self.world = len(self.devices)
        self.n_total = n_total
        self.reduce = reduce or ("nccl" if self.world > 1 else "host")
        compiled = sim.compile()
        self.shards = [shard(n_total, g, self.world) for g in range(self.world)]
        self._pool = ThreadPoolExecutor(max_workers=self.world, thread_name_prefix="itsb-gpu")

        def make(g: int) -> Any:
            first, count = self.shards[g]
            eng = Engine(compiled, device=self.devices[g], lib=self._lib)
            if count:
                eng.init_replicas(count, seed=seed, first_replica_id=first_replica_id + first)
            return eng

        self.engines: List[Any] = self._each(make)
        self._comms: List[Optional[C.c_void_p]] = [None] * self.world
        if self.reduce == "nccl" and self.world > 1:
            uid = np.zeros(128, dtype=np.uint8)
            rc = self._lib.itsb_nccl_unique_id(uid.ctypes.data, 128)
            if rc != 0:
                raise L.EngineError(rc, "itsb_nccl_unique_id")

            def join(g: int) -> C.c_void_p:
                comm = C.c_void_p()
                eng = self.engines[g]
                eng._check(self._lib.itsb_nccl_comm_init(eng._h, self.world, g, uid.ctypes.data, C.byref(comm)),
                           "itsb_nccl_comm_init")
                return comm

            self._comms = self._each(join)          # ncclCommInitRank returns once every rank has joined

    # -- plumbing ---------------------------------------------------------------------------------------------------
    def _each(self, fn: Callable[[int], Any]) -> List[Any]:
        """``fn(g)`` for every shard, concurrently (one thread per GPU); re-raises the first failure."""
        return list(self._pool.map(fn, range(self.world)))

    def _active(self, g: int) -> bool:
        return self.shards[g][1] > 0

    # -- Simulator.run over the whole set -------------------------------------------------------------------------------
    def run(self, duration: float) -> int:
        """``Simulator.run(duration)`` on every replica of every shard; returns the simulated events processed."""
        return sum(self._each(lambda g: self.engines[g].run(duration) if self._active(g) else 0))

    def run_async(self, duration: float) -> None:
        self._each(lambda g: self.engines[g].run_async(duration) if self._active(g) else None)

    def sync(self) -> int:
        return sum(self._each(lambda g: self.engines[g].sync() if self._active(g) else 0))

    def last_run_ms(self) -> float:
        """Device time of the last completed run: the slowest shard's (they run side by side)."""
        return max(self.engines[g].last_run_ms() for g in range(self.world) if self._active(g))

    def shard_ms(self) -> List[float]:
        return [self.engines[g].last_run_ms() if self._active(g) else 0.0 for g in range(self.world)]

    def kernel_names(self) -> List[str]:
        return [e.kernel_name() for e in self.engines]

    def launch_count(self) -> int:
        return sum(e.launch_count() for e in self.engines)

    # -- the one collective ------------------------------------------------------------------------------------------------
    def telemetry(self) -> np.ndarray:
        """The telemetry vector summed over every replica of the job."""
        if self.reduce == "nccl" and self.world > 1:
            return self._each(lambda g: self.engines[g].allreduce_telemetry(self._comms[g]))[0]
        parts = self._each(lambda g: self.engines[g].telemetry() if self._active(g) else np.zeros(L.TM_SIZE, np.uint64))
        return np.sum(np.array(parts, dtype=np.uint64), axis=0, dtype=np.uint64)

    def histogram(self, slot: int, lo: int, width: int, nbins: int, only_nonzero: bool = False) -> np.ndarray:
        """Histogram over ALL replicas of one telemetry slot (config 3's per-replica statistics)."""
        if self.reduce == "nccl" and self.world > 1:
            return self._each(lambda g: self.engines[g].histogram(slot, lo, width, nbins, only_nonzero,
                                                                  comm=self._comms[g]))[0]
        parts = self._each(lambda g: self.engines[g].histogram(slot, lo, width, nbins, only_nonzero)
                           if self._active(g) else np.zeros(nbins, np.uint64))
        return np.sum(np.array(parts, dtype=np.uint64), axis=0, dtype=np.uint64)

    # -- per-replica reads: global replica index -> (shard, local index) ------------------------------------------------
    def locate(self, replica: int):
        for g, (first, count) in enumerate(self.shards):
            if first <= replica < first + count:
                return self.engines[g], replica - first
        raise IndexError(replica)

    def replica_telemetry(self, replica: int) -> np.ndarray:
        eng, r = self.locate(replica)
        return eng.replica_telemetry(r)

    def now(self) -> np.ndarray:
        return np.concatenate([self.engines[g].now() for g in range(self.world) if self._active(g)])

    def close(self) -> None:
        for g, comm in enumerate(self._comms):
            if comm:
                self._lib.itsb_nccl_comm_destroy(comm)
        self._comms = [None] * self.world
        for e in self.engines:
            e.close()
        self._pool.shutdown(wait=True)

    def __enter__(self) -> "ShardedEngine":
        return self

    def __exit__(self, *exc: Any) -> None:
        self.close()


def run_sharded(sim: Any, n_total: int, seed: int = 0, until: Optional[float] = None,
                devices: Optional[Sequence[int]] = None, **kwargs: Any) -> ShardedEngine:
    """Builds the shards and, when ``until`` is given, advances every replica to that simulated time."""
    eng = ShardedEngine(sim, n_total, seed=seed, devices=devices, **kwargs)
    if until is not None:
        eng.run(until)
    return eng


def device_count() -> int:
    """GPUs the CUDA runtime shows (``cudaGetDeviceCount`` through libcudart; 0 when there is none)."""
    for name in ("libcudart.so", "libcudart.so.12", "libcudart.so.13"):
        try:
            rt = C.CDLL(name)
        except OSError:
            continue
        n = C.c_int(0)
        if rt.cudaGetDeviceCount(C.byref(n)) == 0:
            return n.value
        return 0
    return 0


# ---- one process per GPU under an external launcher --------------------------------------------------------------------
class RankEnv:
    """``RANK`` / ``LOCAL_RANK`` / ``WORLD_SIZE`` as a launcher such as torchrun exports them (single node), plus the
    three things a rank needs from the others: the NCCL id, a barrier, and max / sum of a number."""

    def __init__(self, env: Optional[dict] = None, directory: Optional[str] = None) -> None:
        env = os.environ if env is None else env
        self.rank = int(env.get("RANK", "0"))
        self.local_rank = int(env.get("LOCAL_RANK", str(self.rank)))
        self.world = int(env.get("WORLD_SIZE", "1"))
        # every rank of one job sees the same launcher process and the same master port
        job = f"{env.get('MASTER_PORT', '0')}_{env.get('TORCHELASTIC_RUN_ID', 'none')}_{os.getppid()}"
        self._path = os.path.join(directory or tempfile.gettempdir(), f"itsb_nccl_id_{job}")
        self._t0 = time.time()
        self.comm: Optional[C.c_void_p] = None
        self._engine: Any = None

    def first_and_count(self, n_total: int):
        return shard(n_total, self.rank, self.world)

    def exchange(self, payload: Optional[bytes], size: int, timeout: float = 120.0) -> bytes:
        """Rank 0 publishes ``payload`` (``size`` bytes), every rank returns it.  File in the temp directory, written
        under another name and renamed, so a reader never sees half of it; a file older than this job is ignored."""
        if self.world == 1:
            return payload or b""
        if self.rank == 0:
            tmp = self._path + f".{os.getpid()}"
            with open(tmp, "wb") as fh:
                fh.write(struct.pack("<d", time.time()) + payload)
            os.replace(tmp, self._path)
            return payload
        deadline = time.time() + timeout
        while time.time() < deadline:
            try:
                with open(self._path, "rb") as fh:
                    blob = fh.read()
                if len(blob) == 8 + size and struct.unpack("<d", blob[:8])[0] >= self._t0 - 60.0:
                    return blob[8:]
            except OSError:
                pass
            time.sleep(0.01)
        raise TimeoutError(f"rank {self.rank}: no NCCL id from rank 0 at {self._path}")

    def join(self, engine: Any) -> None:
        """Creates this rank's end of the communicator over all ranks (``itsb_nccl_*``)."""
        self._engine = engine
        if self.world == 1:
            return
        lib = engine._lib
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            rc = lib.itsb_nccl_unique_id(uid.ctypes.data, 128)
            if rc != 0:
                raise L.EngineError(rc, "itsb_nccl_unique_id")
        uid = np.frombuffer(self.exchange(uid.tobytes() if self.rank == 0 else None, 128), dtype=np.uint8).copy()
        comm = C.c_void_p()
        engine._check(lib.itsb_nccl_comm_init(engine._h, self.world, self.rank, uid.ctypes.data, C.byref(comm)),
                      "itsb_nccl_comm_init")
        self.comm = comm

    def allreduce(self, values: Sequence[int], op: int = NCCL_SUM) -> np.ndarray:
        """uint64 all-reduce over the ranks (``itsb_allreduce_u64``); the identity with one rank."""
        buf = np.array(values, dtype=np.uint64)
        if self.world > 1:
            eng = self._engine
            eng._check(eng._lib.itsb_allreduce_u64(eng._h, self.comm, buf.ctypes.data, len(buf), op), "itsb_allreduce_u64")
        return buf

    def barrier(self) -> None:
        self.allreduce([1])

    def max_float(self, x: float) -> float:
        """max over the ranks of a non-negative float (order of non-negative doubles = order of their bit patterns)."""
        bits = struct.unpack("<Q", struct.pack("<d", float(x)))[0]
        return struct.unpack("<d", struct.pack("<Q", int(self.allreduce([bits], NCCL_MAX)[0])))[0]

    def sum_int(self, x: int) -> int:
        return int(self.allreduce([int(x)])[0])

    def close(self) -> None:
        joined = bool(self.comm)
        if self.comm:
            self._engine._lib.itsb_nccl_comm_destroy(self.comm)
            self.comm = None
        if self.rank == 0 and joined:       # every rank has read the id: creating the communicator waited for them
            try:
                os.remove(self._path)
            except OSError:
                pass
