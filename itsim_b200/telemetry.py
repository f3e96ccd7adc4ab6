"""
Device telemetry -> the reference datastore's formats.

The reference stores telemetry as JSON documents validated against ``itsim/schemas/items.py`` in SQLite tables
``(uuid, timestamp, sim_uuid, json)`` (``itsim/datastore/database.py:44-57``), one HTTP POST and one SQLite connection
per record (``itsim/logging.py:27-37``).  The engine writes fixed 32-byte ``network_event`` records to HBM
(``itsb_net_event`` in include/itsb.h); this module converts a batch of them to the reference's JSON schema
(``NETWORK_EVENT_SCHEMA``, items.py:83-131) and to rows ready for one bulk ``executemany`` (database.py:130-147).

Two forms of every converter:

* ``*_documents`` -- one Python dict per record, for a handful of records (tests, notebooks);
* ``*_table`` / ``*_jsonl`` -- whole structured arrays at once (what ``Engine.drain_end`` returns: millions of records
  per step), columns built with numpy string kernels, no per-record Python: ids are counter-based UUIDs (a 64-bit
  prefix per simulation and table, a 64-bit record counter) instead of one ``uuid4()`` call per record, addresses are
  formatted once per distinct value.  The same three tables the datastore keeps -- ``network_event``, ``node``, ``log``
  (items.py:34-131, database.py:44-57) -- plus the connection lines of ``Socket.log_pair``.
"""
import json
import uuid
from ipaddress import IPv4Address
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

NETWORK_EVENT_TYPES = {1: "open", 2: "close"}          # items.py:78-81
PROTOCOLS = {0: "NONE", 1: "UDP", 2: "TCP"}            # items.py:106-109


def network_event_documents(records: np.ndarray, sim_uuid: Optional[str] = None,
                            node_uuids: Optional[Sequence[str]] = None) -> List[Dict[str, Any]]:
    """One dict per record with exactly the properties of ``NETWORK_EVENT_SCHEMA``.  ``timestamp`` is the simulated
    time (the reference stamps wall-clock time at logging; the simulation itself has no notion of it)."""
    sim_uuid = sim_uuid or str(uuid.uuid4())
    docs = []
    for rec in records:
        node = int(rec["node"])
        docs.append({
            "sim_uuid": sim_uuid,
            "timestamp": f"{float(rec['t']):.9f}",
            "type": "network_event",
            "uuid": str(uuid.uuid4()),
            "uuid_node": node_uuids[node] if node_uuids is not None else str(uuid.UUID(int=node)),
            "network_event_type": NETWORK_EVENT_TYPES[int(rec["type"])],
            "pid": int(rec["pid"]),
            "protocol": PROTOCOLS[int(rec["protocol"])],
            "src": [str(IPv4Address(int(rec["src_ip"]))), int(rec["port"])],
            "dst": [str(IPv4Address(int(rec["dst_ip"]))), int(rec["dst_port"])],
        })
    return docs


def network_event_labels(records: np.ndarray) -> List[List[str]]:
    """The tag names (``itsim.Tag``) of the process behind each record: the ground truth a detector trained on the
    telemetry is scored against (``sphinx/source/overview.rst:46-56``: tags cascade from malware to what it causes)."""
    from .tags import tag_names
    return [tag_names(int(rec["tags"])) for rec in records]


def datastore_rows(documents: Iterable[Dict[str, Any]]) -> List[Tuple[str, str, str, str]]:
    """``(uuid, timestamp, sim_uuid, json)`` rows of the datastore's ``network_event`` table."""
    return [(d["uuid"], d["timestamp"], d["sim_uuid"], json.dumps(d)) for d in documents]


def connection_documents(records: np.ndarray) -> List[Dict[str, Any]]:
    """The dictionaries ``Socket.log_pair`` dumps as JSON log lines (network_simple_overrides.py:163-177), value
    conventions included: a missing half shows as integer 0 (``Remote_IP`` as the string "0"), a present one as
    strings."""
    docs = []
    for rec in records:
        has_in, has_out = bool(int(rec["has"]) & 1), bool(int(rec["has"]) & 2)
        start = float(rec["t_start"])
        docs.append({
            "Connection_Type": "UDP",
            "Local_IP": str(IPv4Address(int(rec["local_ip"]))) if has_in else 0,
            "Local_Port": str(int(rec["local_port"])) if has_in else 0,
            "Direction": "Inbound" if int(rec["inbound"]) else "Outbound",
            "Remote_IP": str(IPv4Address(int(rec["remote_ip"]))) if has_out else "0",
            "Remote_Port": str(int(rec["remote_port"])) if has_out else 0,
            "Sent_Bytes": str(int(rec["sent_bytes"])) if has_out else 0,
            "Received_Bytes": str(int(rec["received_bytes"])) if has_in else 0,
            "PID": 0,
            "Start_Time": start,
            "End_Time": float(rec["t_end"]),
            "Parent_Process_Path": "/home/town",
            "Parent_Process_Start_Time": start,
        })
    return docs


# ---- bulk converters: structured arrays in, columns of strings out ---------------------------------------------------------
def _uuid_column(prefix64: int, first: int, n: int) -> np.ndarray:
    """``n`` UUID strings ``pppppppp-pppp-pppp-cccc-cccccccccccc``: 64-bit prefix, 64-bit counter from ``first``."""
    head = f"{(prefix64 >> 32) & 0xFFFFFFFF:08x}-{(prefix64 >> 16) & 0xFFFF:04x}-{prefix64 & 0xFFFF:04x}-"
    c = np.arange(first, first + n, dtype=np.uint64)
    hi = np.char.mod("%04x", (c >> np.uint64(48)).astype(np.int64))
    lo = np.char.mod("%012x", (c & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64))
    return np.char.add(np.char.add(np.char.add(head, hi), "-"), lo)


def _prefix(sim_uuid: str, table: str) -> int:
    """64-bit id prefix of one (simulation, table): stable, distinct per table."""
    import hashlib
    return int.from_bytes(hashlib.sha256(f"{sim_uuid}/{table}".encode()).digest()[:8], "big")


def _ip_column(values: np.ndarray) -> np.ndarray:
    """Dotted-quad strings of an array of IPv4 addresses, each distinct value formatted once."""
    uniq, inverse = np.unique(values, return_inverse=True)
    return np.array([str(IPv4Address(int(v))) for v in uniq], dtype=object)[inverse].astype(str)


def _cat(*parts: Any) -> np.ndarray:
    out = parts[0]
    for p in parts[1:]:
        out = np.char.add(out, p)
    return out


def _replica_column(first: np.ndarray, n_records: int) -> np.ndarray:
    """Replica index of every record of a drain, from its ``n_replicas + 1`` offsets."""
    counts = np.diff(first.astype(np.int64))
    return np.repeat(np.arange(len(counts), dtype=np.int64), counts)[:n_records]


def network_event_table(records: np.ndarray, sim_uuid: str, first_record: int = 0,
                        replica_of: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """Columns ``uuid, timestamp, sim_uuid, json`` of the datastore's ``network_event`` table (database.py:44-57) for a
    whole array of device records.  ``uuid_node`` encodes (replica, node): the node's index in the low 32 bits, the
    replica (``replica_of``, e.g. from a drain's ``net_first``) above it -- replicas are separate worlds."""
    n = len(records)
    uuids = _uuid_column(_prefix(sim_uuid, "network_event"), first_record, n)
    stamps = np.char.mod("%.9f", records["t"])
    rep = replica_of if replica_of is not None else np.zeros(n, dtype=np.int64)
    node_id = (rep.astype(np.uint64) << np.uint64(32)) | records["node"].astype(np.uint64)
    node_uuid = _uuid_column(_prefix(sim_uuid, "node"), 0, 0) if n == 0 else \
        _cat(f"{(_prefix(sim_uuid, 'node') >> 32) & 0xFFFFFFFF:08x}-{(_prefix(sim_uuid, 'node') >> 16) & 0xFFFF:04x}-"
             f"{_prefix(sim_uuid, 'node') & 0xFFFF:04x}-",
             np.char.mod("%04x", (node_id >> np.uint64(48)).astype(np.int64)), "-",
             np.char.mod("%012x", (node_id & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)))
    kinds = np.array(["", "open", "close"], dtype=object)[np.clip(records["type"], 0, 2)].astype(str)
    protos = np.array(["NONE", "UDP", "TCP"], dtype=object)[np.clip(records["protocol"], 0, 2)].astype(str)
    doc = _cat('{"sim_uuid": "' + sim_uuid + '", "timestamp": "', stamps, '", "type": "network_event", "uuid": "', uuids,
               '", "uuid_node": "', node_uuid, '", "network_event_type": "', kinds,
               '", "pid": ', np.char.mod("%d", records["pid"]), ', "protocol": "', protos,
               '", "src": ["', _ip_column(records["src_ip"]), '", ', np.char.mod("%d", records["port"]),
               '], "dst": ["', _ip_column(records["dst_ip"]), '", ', np.char.mod("%d", records["dst_port"]), ']}')
    return {"uuid": uuids, "timestamp": stamps, "sim_uuid": np.full(n, sim_uuid), "json": doc}


def node_table(sim: Any, sim_uuid: str, replicas: int = 1, timestamp: str = "0.000000000") -> Dict[str, np.ndarray]:
    """The datastore's ``node`` table (NODE_SCHEMA, items.py:34-51): one document per (replica, node) of the model, with
    the ``uuid`` the ``network_event`` documents refer to as ``uuid_node`` and the node's name as ``node_label``."""
    labels = np.array([getattr(node, "name", f"node-{i}") for i, node in enumerate(sim.nodes)], dtype=str)
    m = len(labels)
    rep = np.repeat(np.arange(replicas, dtype=np.uint64), m)
    idx = np.tile(np.arange(m, dtype=np.uint64), replicas)
    node_id = (rep << np.uint64(32)) | idx
    pre = _prefix(sim_uuid, "node")
    uuids = _cat(f"{(pre >> 32) & 0xFFFFFFFF:08x}-{(pre >> 16) & 0xFFFF:04x}-{pre & 0xFFFF:04x}-",
                 np.char.mod("%04x", (node_id >> np.uint64(48)).astype(np.int64)), "-",
                 np.char.mod("%012x", (node_id & np.uint64(0xFFFFFFFFFFFF)).astype(np.int64)))
    lab = np.tile(labels, replicas)
    if replicas > 1:
        lab = _cat(lab, " #", np.char.mod("%d", rep.astype(np.int64)))
    doc = _cat('{"sim_uuid": "' + sim_uuid + '", "timestamp": "' + timestamp + '", "type": "node", "uuid": "', uuids,
               '", "node_label": ', np.array([json.dumps(x) for x in lab], dtype=str), '}')
    return {"uuid": uuids, "timestamp": np.full(len(uuids), timestamp), "sim_uuid": np.full(len(uuids), sim_uuid), "json": doc}


LOG_KINDS = {1: "process started", 2: "advance() returned", 3: "transmission started", 4: "transmission ended",
             5: "packet delivered", 6: "process resumed", 7: "exception delivered", 8: "run() horizon reached"}


def log_table(trace: np.ndarray, sim_uuid: str, program_names: Optional[Dict[int, str]] = None,
              first_record: int = 0, level: str = "INFO") -> Dict[str, np.ndarray]:
    """The datastore's ``log`` table (LOG_SCHEMA, items.py:54-75) from an engine trace (``Engine.trace``): one line per
    simulated event -- what the reference's example loggers print per event (``network_simple.py:26-34``: simulated
    time, process, message), with the levels of ``README.md:18-50``."""
    n = len(trace)
    uuids = _uuid_column(_prefix(sim_uuid, "log"), first_record, n)
    stamps = np.char.mod("%.9f", trace["t"])
    names = program_names or {}
    kind_txt = np.array([LOG_KINDS.get(k, f"event {k}") for k in range(256)], dtype=object)[trace["kind"]].astype(str)
    prog_txt = np.array([names.get(k, f"program {k}") for k in range(256)], dtype=object)[trace["prog"]].astype(str)
    content = _cat(kind_txt, ": ", prog_txt, " on node ", np.char.mod("%d", trace["node"]))
    doc = _cat('{"sim_uuid": "' + sim_uuid + '", "timestamp": "', stamps, '", "type": "log", "uuid": "', uuids,
               '", "content": "', content, '", "level": "' + level + '"}')
    return {"uuid": uuids, "timestamp": stamps, "sim_uuid": np.full(n, sim_uuid), "json": doc}


def connection_jsonl(records: np.ndarray) -> np.ndarray:
    """The JSON lines ``Socket.log_pair`` prints (network_simple_overrides.py:163-177), one string per record, value
    conventions as in :func:`connection_documents` (which this equals record for record after ``json.loads``)."""
    n = len(records)
    has_in = (records["has"] & 1) != 0
    has_out = (records["has"] & 2) != 0

    def quoted_or_zero(present: np.ndarray, text: np.ndarray) -> np.ndarray:
        return np.where(present, _cat('"', text, '"'), "0")

    local_ip = quoted_or_zero(has_in, _ip_column(records["local_ip"]))
    local_port = quoted_or_zero(has_in, np.char.mod("%d", records["local_port"]))
    remote_ip = np.where(has_out, _cat('"', _ip_column(records["remote_ip"]), '"'), '"0"')
    remote_port = quoted_or_zero(has_out, np.char.mod("%d", records["remote_port"]))
    sent = quoted_or_zero(has_out, np.char.mod("%d", records["sent_bytes"]))
    received = quoted_or_zero(has_in, np.char.mod("%d", records["received_bytes"]))
    direction = np.where(records["inbound"] != 0, '"Inbound"', '"Outbound"')
    start = np.array([repr(float(x)) for x in records["t_start"]], dtype=str) if n < 4096 else np.char.mod("%.17g", records["t_start"])
    end = np.array([repr(float(x)) for x in records["t_end"]], dtype=str) if n < 4096 else np.char.mod("%.17g", records["t_end"])
    return _cat('{"Connection_Type": "UDP", "Local_IP": ', local_ip, ', "Local_Port": ', local_port, ', "Direction": ',
                direction, ', "Remote_IP": ', remote_ip, ', "Remote_Port": ', remote_port, ', "Sent_Bytes": ', sent,
                ', "Received_Bytes": ', received, ', "PID": 0, "Start_Time": ', start, ', "End_Time": ', end,
                ', "Parent_Process_Path": "/home/town", "Parent_Process_Start_Time": ', start, '}')


def table_rows(table: Dict[str, np.ndarray]) -> List[Tuple[str, str, str, str]]:
    """``(uuid, timestamp, sim_uuid, json)`` tuples for one ``executemany`` (database.py:130-147)."""
    return list(zip(table["uuid"].tolist(), table["timestamp"].tolist(), table["sim_uuid"].tolist(), table["json"].tolist()))
