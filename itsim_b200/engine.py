"""
Engine: one batch of replicas of one compiled model on one GPU, driven through the C ABI.

Mirrors the reference's run-time surface -- ``Simulator.run(duration)`` / ``Simulator.now()`` (greensim, subclassed
at ``itsim/simulator.py:9``; repeated ``run`` continues, ``tests/machine/process_management/test_thread.py:148-152``)
-- for N replicas at once.
"""
import ctypes as C
from typing import Any, Dict, Optional

import numpy as np

from . import _lib as L


class Engine:

    def __init__(self, compiled: "Any", device: int = 0, lib: Optional[C.CDLL] = None) -> None:
        self._lib = lib if lib is not None else L.load()
        self._h = C.c_void_p()
        self._check(self._lib.itsb_create(device, C.byref(self._h)), "itsb_create")
        self.compiled = compiled
        blobs = compiled.blobs()
        ptrs = (C.c_void_p * len(blobs))(*[b.ctypes.data if b.size else None for b in blobs])
        sizes = (C.c_size_t * len(blobs))(*[b.nbytes for b in blobs])
        self._check(self._lib.itsb_load_model(self._h, C.byref(compiled.desc()), ptrs, sizes), "itsb_load_model")
        self.n_replicas = 0

    # -- plumbing ----------------------------------------------------------------------------------------
    def _check(self, rc: int, where: str) -> None:
        if rc != 0:
            detail = ""
            if self._h:
                msg = self._lib.itsb_last_error(self._h)
                detail = msg.decode() if msg else ""
                bad = self._lib.itsb_first_failed_replica(self._h) if self.__dict__.get("n_replicas") else -1
                if bad >= 0:
                    st, dt = C.c_int32(), C.c_uint64()
                    self._lib.itsb_replica_status(self._h, bad, C.byref(st), C.byref(dt))
                    detail += f" replica {bad}: status {st.value}, detail {dt.value:#x}"
            raise L.EngineError(rc, where, detail)

    def close(self) -> None:
        if self._h:
            self._lib.itsb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    # -- replicas ------------------------------------------------------------------------------------------
    def init_replicas(self, n: int, seed: int = 0, first_replica_id: int = 0) -> None:
        self._check(self._lib.itsb_init_replicas(self._h, n, seed, first_replica_id), "itsb_init_replicas")
        self.n_replicas = n

    def run(self, duration: float) -> int:
        """``Simulator.run(duration)`` on every replica; returns simulated events processed by this call."""
        done = C.c_uint64()
        self._check(self._lib.itsb_run(self._h, float(duration), C.byref(done)), "itsb_run")
        return done.value

    def run_async(self, duration: float) -> None:
        self._check(self._lib.itsb_run_async(self._h, float(duration)), "itsb_run_async")

    def sync(self) -> int:
        done = C.c_uint64()
        self._check(self._lib.itsb_sync(self._h, C.byref(done)), "itsb_sync")
        return done.value

    def last_run_ms(self) -> float:
        ms = C.c_float()
        self._check(self._lib.itsb_last_run_ms(self._h, C.byref(ms)), "itsb_last_run_ms")
        return ms.value

    def kernel_name(self) -> str:
        """The engine kernel ``run`` launches for this model (``itsb_kernel_name``)."""
        name = self._lib.itsb_kernel_name(self._h)
        return name.decode() if name else ""

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.itsb_launch_count(self._h, C.byref(n)), "itsb_launch_count")
        return n.value

    def now(self, first: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = self.n_replicas - first if n is None else n
        out = np.zeros(n, dtype=np.float64)
        self._check(self._lib.itsb_read_now(self._h, first, n, out.ctypes.data), "itsb_read_now")
        return out

    def trace(self, replica: int = 0) -> np.ndarray:
        cap = self.compiled.caps.trace_records
        out = np.zeros(cap, dtype=L.TRACE_DTYPE)
        total = C.c_size_t()
        self._check(self._lib.itsb_read_trace(self._h, replica, out.ctypes.data, cap, C.byref(total)), "itsb_read_trace")
        self.trace_total = total.value
        return out[:min(cap, total.value)]

    def net_events(self, replica: int = 0) -> np.ndarray:
        """The replica's ``network_event`` log (sockets opened / closed), as device records."""
        cap = self.compiled.caps.net_event_records
        out = np.zeros(cap, dtype=L.NET_EVENT_DTYPE)
        total = C.c_size_t()
        self._check(self._lib.itsb_read_net_events(self._h, replica, out.ctypes.data, cap, C.byref(total)),
                    "itsb_read_net_events")
        self.net_events_total = total.value
        return out[:min(cap, total.value)]

    def connections(self, replica: int = 0) -> np.ndarray:
        """The replica's connection records (``Socket.log_pair``, network_simple_overrides.py:152-177)."""
        cap = self.compiled.caps.connection_records
        out = np.zeros(cap, dtype=L.CONNECTION_DTYPE)
        total = C.c_size_t()
        self._check(self._lib.itsb_read_connections(self._h, replica, out.ctypes.data, cap, C.byref(total)),
                    "itsb_read_connections")
        self.connections_total = total.value
        return out[:min(cap, total.value)]

    # -- streaming the record logs home (itsb_drain_begin / itsb_drain_end) ----------------------------------------------
    def drain_begin(self) -> None:
        """Packs every replica's new ``network_event`` / connection records on the device, empties the per-replica logs
        and starts the copy to pinned host memory on a second stream; the next ``run`` may be queued right away."""
        self._check(self._lib.itsb_drain_begin(self._h), "itsb_drain_begin")

    def drain_end(self, copy: bool = False) -> Dict[str, Any]:
        """Waits for the oldest pending drain.  Returns ``net`` / ``conn`` (structured arrays over ALL replicas,
        replica after replica), ``net_first`` / ``conn_first`` (``n_replicas + 1`` offsets: replica r owns
        ``[first[r], first[r + 1])``), the records lost to full per-replica logs, and the bytes copied.  The arrays
        are views of the engine's pinned buffers, valid until the drain after the next one begins (``copy=True``
        detaches them)."""
        d = L.Drain()
        self._check(self._lib.itsb_drain_end(self._h, C.byref(d)), "itsb_drain_end")
        n = self.n_replicas

        def view(ptr: Any, count: int, dtype: Any) -> np.ndarray:
            if not ptr or count == 0:
                return np.zeros(0, dtype=dtype)
            buf = (C.c_char * (count * np.dtype(dtype).itemsize)).from_address(ptr)
            arr = np.frombuffer(buf, dtype=dtype, count=count)
            return arr.copy() if copy else arr

        return {"net": view(d.net, d.n_net, L.NET_EVENT_DTYPE), "net_first": view(d.net_first, n + 1, np.uint64),
                "conn": view(d.conn, d.n_conn, L.CONNECTION_DTYPE), "conn_first": view(d.conn_first, n + 1, np.uint64),
                "net_dropped": int(d.net_dropped), "conn_dropped": int(d.conn_dropped), "d2h_bytes": int(d.d2h_bytes),
                "pack_ms": float(d.pack_ms)}

    def drain(self, copy: bool = False) -> Dict[str, Any]:
        self.drain_begin()
        return self.drain_end(copy=copy)

    def histogram(self, slot: int, lo: int, width: int, nbins: int, only_nonzero: bool = False,
                  comm: Optional[C.c_void_p] = None) -> np.ndarray:
        """Histogram over the replicas of telemetry slot ``slot`` (``_lib.tm_user(k)`` for model counter k); with a
        communicator the bins are summed over the ranks (one ncclAllReduce of uint64 bins)."""
        out = np.zeros(nbins, dtype=np.uint64)
        self._check(self._lib.itsb_histogram(self._h, slot, lo, width, nbins, 1 if only_nonzero else 0, comm,
                                             out.ctypes.data), "itsb_histogram")
        return out

    def telemetry(self) -> np.ndarray:
        out = np.zeros(L.TM_SIZE, dtype=np.uint64)
        self._check(self._lib.itsb_read_telemetry(self._h, out.ctypes.data, L.TM_SIZE), "itsb_read_telemetry")
        return out

    def replica_telemetry(self, replica: int) -> np.ndarray:
        out = np.zeros(L.TM_SIZE, dtype=np.uint64)
        self._check(self._lib.itsb_read_replica_telemetry(self._h, replica, out.ctypes.data, L.TM_SIZE),
                    "itsb_read_replica_telemetry")
        return out

    def allreduce_telemetry(self, comm: C.c_void_p) -> np.ndarray:
        out = np.zeros(L.TM_SIZE, dtype=np.uint64)
        self._check(self._lib.itsb_allreduce_telemetry(self._h, comm, out.ctypes.data, L.TM_SIZE),
                    "itsb_allreduce_telemetry")
        return out

    def state_bytes(self) -> Dict[str, int]:
        hot, arena = C.c_uint64(), C.c_uint64()
        self._check(self._lib.itsb_state_bytes(self._h, C.byref(hot), C.byref(arena)), "itsb_state_bytes")
        return {"hot": hot.value, "arena": arena.value}


def summarize(tm: np.ndarray) -> Dict[str, Any]:
    """Names the slots of a telemetry vector."""
    return {
        "events": int(tm[L.TM_KIND0:L.TM_KIND0 + L.NUM_KINDS].sum()),
        "counts": {L.KIND_NAMES[k]: int(tm[L.TM_KIND0 + k]) for k in range(1, L.NUM_KINDS)},
        "packets_sent": int(tm[L.TM_PACKETS_SENT]), "bytes_sent": int(tm[L.TM_BYTES_SENT]),
        "packets_received": int(tm[L.TM_PACKETS_RECEIVED]), "bytes_received": int(tm[L.TM_BYTES_RECEIVED]),
        "delivered": int(tm[L.TM_DELIVERED]), "dropped": int(tm[L.TM_DROPPED]),
        "user": [int(tm[L.tm_user(k)]) for k in range(L.TM_USER_SLOTS)],
        "latency_hist": [int(v) for v in tm[L.TM_LAT0:L.TM_LAT0 + L.LAT_BINS]],
    }
