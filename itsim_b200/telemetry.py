"""
Device telemetry -> the reference datastore's formats.

The reference stores telemetry as JSON documents validated against ``itsim/schemas/items.py`` in SQLite tables
``(uuid, timestamp, sim_uuid, json)`` (``itsim/datastore/database.py:44-57``), one HTTP POST and one SQLite connection
per record (``itsim/logging.py:27-37``).  The engine writes fixed 32-byte ``network_event`` records to HBM
(``itsb_net_event`` in include/itsb.h); this module converts a batch of them to the reference's JSON schema
(``NETWORK_EVENT_SCHEMA``, items.py:83-131) and to rows ready for one bulk ``executemany`` (database.py:130-147).
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
