"""
ctypes binding of ``libitsim_b200.so`` (C ABI in ``include/itsb.h``).

This is the only way the package reaches compute: there is **no CPU execution path**.  If the CUDA library has not
been built (``python -c "import __graft_entry__ as g; g.build()"``) importing an engine raises immediately, and
without a CUDA device ``itsb_create`` returns ``ITSB_ERR_NO_DEVICE``.
"""
import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# ITSB_LIB: another build of the same library (the -DITSB_EXPERIMENTS one, for A/B measurements); never a different engine
LIB_PATH = os.environ.get("ITSB_LIB") or os.path.join(HERE, "libitsim_b200.so")

ABI_VERSION = 5
TM_KIND0, TM_PACKETS_SENT, TM_BYTES_SENT, TM_PACKETS_RECEIVED, TM_BYTES_RECEIVED = 0, 9, 10, 11, 12
TM_DELIVERED, TM_DROPPED, TM_USER0, TM_LAT0, LAT_BINS = 13, 14, 15, 23, 64
TM_USER8, TM_USER_SLOTS = TM_LAT0 + LAT_BINS, 12
TM_SIZE = TM_USER8 + 4


def tm_user(slot: int) -> int:
    """Index of model counter ``slot`` in the telemetry vector (ITSB_TM_USER in include/itsb.h)."""
    return TM_USER0 + slot if slot < 8 else TM_USER8 + slot - 8
NUM_KINDS = 9
EV_PROC_START, EV_TIMER, EV_TX_BEGIN, EV_TX_END, EV_DELIVER, EV_WAKE, EV_THROW, EV_STOP = range(1, 9)
TRACE_MARK = 9
KIND_NAMES = {1: "PROC_START", 2: "TIMER", 3: "TX_BEGIN", 4: "TX_END", 5: "DELIVER", 6: "WAKE", 7: "THROW", 8: "STOP"}

ERRORS = {
    -1: "no CUDA device (the engine has no CPU fallback)", -2: "bad model", -3: "CUDA error", -4: "bad call order",
    -10: "event pool exhausted", -11: "process slots exhausted", -12: "socket slots exhausted",
    -13: "in-flight packet slots exhausted", -14: "queue arena exhausted", -15: "bad opcode",
    -16: "thread / process / signal table exhausted", -20: "NCCL error",
    1: "NoSuchAddress / CannotForward", 2: "NoRouteToHost", 3: "PortAlreadyInUse", 4: "EphemeralPortsAllInUse",
    5: "ValueError: Socket is closed", 6: "ValueError: negative delay", 7: "uncaught exception in a process body",
}

NET_EVENT_DTYPE = np.dtype([("t", "<f8"), ("node", "<u2"), ("port", "<u2"), ("type", "u1"), ("protocol", "u1"),
                            ("pid", "<i2"), ("src_ip", "<u4"), ("dst_ip", "<u4"), ("dst_port", "<u2"), ("prog", "<u2"),
                            ("tags", "<u2"), ("reserved", "<u2")])
assert NET_EVENT_DTYPE.itemsize == 32
TRACE_DTYPE = np.dtype([("t", "<f8"), ("kind", "u1"), ("prog", "u1"), ("node", "<u2"), ("aux", "<u4")])
DIST_DTYPE = np.dtype([("kind", "<u4"), ("flags", "<u4"), ("p0", "<f8"), ("p1", "<f8"), ("slope", "<f8"),
                       ("shift", "<f8"), ("lower", "<f8"), ("upper", "<f8")])
NET_DTYPE = np.dtype([("cidr_net", "<u4"), ("cidr_mask", "<u4"), ("bcast", "<u4"), ("first_node", "<u4"),
                      ("n_nodes", "<u4"), ("lat_dist", "<u4"), ("bw_dist", "<u4"), ("mode", "<u4")])
NODE_DTYPE = np.dtype([("addr", "<u4"), ("net", "<u4"), ("host", "<u4"), ("flags", "<u4")])
ENTRY_DTYPE = np.dtype([("pc", "<u4"), ("prog", "<u4")])
INIT_DTYPE = np.dtype([("entry", "<u4"), ("node", "<u4"), ("r", "<u4", (4,)), ("mode", "<u4"), ("delay_const", "<u4")])
IFACE_DTYPE = np.dtype([("node", "<u4"), ("net", "<u4"), ("addr0", "<u4"), ("first_route", "<u4"), ("n_routes", "<u4"),
                        ("reserved", "<u4", (3,))])
ROUTE_DTYPE = np.dtype([("net", "<u4"), ("mask", "<u4"), ("prefixlen", "<u4"), ("kind", "<u4"), ("gateway", "<u4"),
                        ("reserved", "<u4", (3,))])
FILTER_DTYPE = np.dtype([("tags_always", "<u4"), ("tags_if_host", "<u4"), ("tags_if_list", "<u4"), ("tags_if_listed", "<u4")])
CONNECTION_DTYPE = np.dtype([("t_end", "<f8"), ("t_start", "<f8"), ("local_ip", "<u4"), ("remote_ip", "<u4"),
                             ("sent_bytes", "<u4"), ("received_bytes", "<u4"), ("local_port", "<u2"),
                             ("remote_port", "<u2"), ("node", "<u2"), ("inbound", "u1"), ("has", "u1")])
assert CONNECTION_DTYPE.itemsize == 40
assert TRACE_DTYPE.itemsize == 16 and DIST_DTYPE.itemsize == 56 and NET_DTYPE.itemsize == 32
assert NODE_DTYPE.itemsize == 16 and ENTRY_DTYPE.itemsize == 8 and INIT_DTYPE.itemsize == 32
assert IFACE_DTYPE.itemsize == 32 and ROUTE_DTYPE.itemsize == 32


class Caps(C.Structure):
    _fields_ = [("max_events", C.c_uint32), ("max_procs", C.c_uint32), ("max_sockets", C.c_uint32),
                ("max_packets", C.c_uint32), ("arena_nodes", C.c_uint32), ("trace_records", C.c_uint32),
                ("max_threads", C.c_uint32), ("max_tasks", C.c_uint32), ("max_signals", C.c_uint32),
                ("table_words", C.c_uint32), ("net_event_records", C.c_uint32),
                ("connection_records", C.c_uint32), ("connection_window", C.c_uint32)]


class ModelDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_code", C.c_uint32), ("n_entries", C.c_uint32),
                ("n_dists", C.c_uint32), ("n_consts", C.c_uint32), ("n_nets", C.c_uint32), ("n_nodes", C.c_uint32),
                ("n_init", C.c_uint32), ("n_filters", C.c_uint32), ("n_ifaces", C.c_uint32), ("n_routes", C.c_uint32),
                ("n_members", C.c_uint32), ("address_tags", C.c_uint32), ("caps", Caps)]


class Drain(C.Structure):
    """``itsb_drain`` (include/itsb.h): one emptying of every replica's record logs."""
    _fields_ = [("net", C.c_void_p), ("n_net", C.c_uint64), ("net_first", C.c_void_p),
                ("conn", C.c_void_p), ("n_conn", C.c_uint64), ("conn_first", C.c_void_p),
                ("net_dropped", C.c_uint64), ("conn_dropped", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("pack_ms", C.c_double)]


class EngineError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = "") -> None:
        self.code = code
        super().__init__(f"{where}: {ERRORS.get(code, 'error')} (code {code}){': ' + detail if detail else ''}")


_SIGNATURES = {
    "itsb_abi_version": (C.c_int, []),
    "itsb_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "itsb_load_model": (C.c_int, [C.c_void_p, C.POINTER(ModelDesc), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "itsb_init_replicas": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]),
    "itsb_run": (C.c_int, [C.c_void_p, C.c_double, C.POINTER(C.c_uint64)]),
    "itsb_run_async": (C.c_int, [C.c_void_p, C.c_double]),
    "itsb_sync": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "itsb_last_run_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "itsb_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "itsb_kernel_name": (C.c_char_p, [C.c_void_p]),
    "itsb_read_now": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "itsb_read_trace": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "itsb_read_net_events": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "itsb_histogram": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p,
                                 C.c_void_p]),
    "itsb_read_connections": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "itsb_drain_begin": (C.c_int, [C.c_void_p]),
    "itsb_drain_end": (C.c_int, [C.c_void_p, C.POINTER(Drain)]),
    "itsb_read_telemetry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsb_read_replica_telemetry": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t]),
    "itsb_allreduce_telemetry": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsb_allreduce_u64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "itsb_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "itsb_nccl_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "itsb_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "itsb_replica_status": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_uint64)]),
    "itsb_first_failed_replica": (C.c_int64, [C.c_void_p]),
    "itsb_state_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "itsb_eval_log": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "itsb_last_error": (C.c_char_p, [C.c_void_p]),
    "itsb_destroy": (None, [C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def bind(lib: C.CDLL) -> C.CDLL:
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.itsb_abi_version() != ABI_VERSION:
        raise ImportError(f"libitsim_b200 ABI {lib.itsb_abi_version()} != binding {ABI_VERSION}")
    return lib


def load(path: Optional[str] = None) -> C.CDLL:
    """Loads the CUDA engine.  Fails loudly when it is missing: nothing else can run a model."""
    global _lib
    if path is None and _lib is not None:
        return _lib
    target = path or LIB_PATH
    if not os.path.exists(target):
        raise ImportError(
            f"{target} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()'). "
            "itsim_b200 has no CPU fallback.")
    lib = bind(C.CDLL(target))
    if path is None:
        _lib = lib
    return lib
