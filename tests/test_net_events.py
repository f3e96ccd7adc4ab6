"""network_event records (sockets opened / closed): engine log in HBM against the oracle, and the conversion to the
reference datastore's JSON schema (itsim/schemas/items.py:83-131)."""
import json
import os
import sqlite3

import numpy as np
import pytest

BACKENDS = [pytest.param("hostemu", id="hostemu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def engine_log(backend, hostemu, seed, replica, horizon, cap=20000):
    from itsim_b200 import netsimple
    sim = netsimple.build()
    sim.caps["net_event_records"] = cap
    sim._lib = hostemu if backend == "hostemu" else None
    eng = sim.run_replicas(2, seed=seed, first_replica_id=replica)
    eng.run(horizon)
    return eng


KERNEL_BACKENDS = BACKENDS + [pytest.param("gpu:vote", id="gpu-k_run_vote", marks=pytest.mark.gpu),
                              pytest.param("gpu:pool", id="gpu-k_run_pool", marks=pytest.mark.gpu)]


@pytest.mark.parametrize("backend", KERNEL_BACKENDS)
def test_log_matches_the_oracle(backend, hostemu, monkeypatch):
    import netsimple as oracle_netsimple
    seed, replica, horizon = 4, 2, 1500.0
    if backend.startswith("gpu"):
        monkeypatch.setenv("ITSB_KERNEL", backend.split(":")[1] if ":" in backend else "warp")
        backend = "gpu"
    eng = engine_log(backend, hostemu, seed, replica, horizon)
    got = eng.net_events(0)
    model = oracle_netsimple.run_replica(seed, replica, horizon, trace=False)
    ref = list(model.net_events)     # before close(): unwinding the parked bodies closes their sockets too
    model.close()
    assert eng.net_events_total == len(ref) == len(got)
    assert [int(x) for x in got["node"]] == [r[1] for r in ref]
    assert [int(x) for x in got["port"]] == [r[2] for r in ref]
    assert [int(x) for x in got["type"]] == [r[3] for r in ref]
    assert np.array_equal(got["t"], np.array([r[0] for r in ref], dtype=np.float64))
    # every socket that is opened and closed is closed after it was opened, on the same node and port
    opened = {}
    for rec in got:
        key = (int(rec["node"]), int(rec["port"]))
        if rec["type"] == 1:
            assert key not in opened
            opened[key] = float(rec["t"])
        else:
            assert float(rec["t"]) >= opened.pop(key)


@pytest.mark.parametrize("backend", BACKENDS)
def test_overflow_is_counted_not_written(backend, hostemu):
    eng = engine_log(backend, hostemu, 0, 0, 1500.0, cap=16)
    got = eng.net_events(0)
    assert len(got) == 16 and eng.net_events_total > 16


def test_conversion_to_the_datastore_schema(hostemu, tmp_path):
    """The documents carry exactly the properties of NETWORK_EVENT_SCHEMA and load into the datastore's table layout
    (uuid, timestamp, sim_uuid, json) with one executemany."""
    from itsim_b200 import telemetry
    eng = engine_log("hostemu", hostemu, 0, 0, 900.0)
    docs = telemetry.network_event_documents(eng.net_events(0), sim_uuid="11111111-2222-3333-4444-555555555555")
    assert docs
    expected = {"sim_uuid", "timestamp", "type", "uuid", "uuid_node", "network_event_type", "pid", "protocol", "src", "dst"}
    for d in docs:
        assert set(d) == expected
        assert d["network_event_type"] in ("open", "close") and d["protocol"] in ("TCP", "UDP", "NONE")
        assert isinstance(d["src"][0], str) and isinstance(d["src"][1], int)
    assert docs[0]["src"][0].startswith("192.168.4.")
    reference_schema = "/root/reference/itsim/schemas/items.py"
    if os.path.exists(reference_schema):
        import jsonschema
        scope = {}
        source = open(reference_schema).read().split("def get_itsim_object_types")[0]
        source = source.replace("import warlock", "")
        exec(compile(source, reference_schema, "exec"), scope)          # the schema dicts only (warlock is not installed)
        for d in docs[:50]:
            jsonschema.validate(d, scope["NETWORK_EVENT_SCHEMA"])
    rows = telemetry.datastore_rows(docs)
    db = sqlite3.connect(str(tmp_path / "itsim.sqlite"))
    db.execute("CREATE TABLE network_event (uuid TEXT PRIMARY KEY, timestamp TEXT, sim_uuid TEXT, json TEXT)")
    db.executemany("INSERT INTO network_event VALUES (?, ?, ?, ?)", rows)
    assert db.execute("SELECT COUNT(*) FROM network_event").fetchone()[0] == len(docs)
    first = json.loads(db.execute("SELECT json FROM network_event ORDER BY timestamp LIMIT 1").fetchone()[0])
    assert first["network_event_type"] == "open"


@pytest.mark.parametrize("backend", BACKENDS)
def test_records_carry_the_tags_greensim_cascades(backend, hostemu):
    """malware_simple.py tags its bodies (`@malware`, `@tagged(Tag.VULNERABLE)`) and greensim hands a process the tags
    of the process that created it (itsim/__init__.py:52-57): the sockets `exfiltrate` opens are labelled MALWARE
    when `call_home` spawned it, MALWARE + VULNERABLE when it came through `sudo_transmit`, and nothing the
    network_simple programs open is labelled."""
    import netsimple as oracle_netsimple
    from itsim_b200 import netsimple
    from itsim_b200.tags import Tag, tag_mask, tag_names
    seed, replica, horizon = 6, 3, 1200.0
    sim = netsimple.build_malware()
    sim.caps["net_event_records"] = 30000
    sim._lib = hostemu if backend == "hostemu" else None
    eng = sim.run_replicas(1, seed=seed, first_replica_id=replica)
    eng.run(horizon)
    got = eng.net_events(0)
    model = oracle_netsimple.MalwareModel(seed, replica, trace=False)
    model.run(horizon)
    ref = list(model.net_events)
    model.close()
    assert len(got) == len(ref)
    assert [int(x) for x in got["type"]] == [r[3] for r in ref]
    assert [int(x) for x in got["port"]] == [r[2] for r in ref]
    assert [int(x) for x in got["tags"]] == [r[4] for r in ref]
    masks = {int(x) for x in got["tags"]}
    assert masks == {0, tag_mask([Tag.MALWARE]), tag_mask([Tag.MALWARE, Tag.VULNERABLE])}
    assert tag_names(3) == ["MALWARE", "VULNERABLE"]
    # only the overlay's own sockets (port 666: exfiltrate and the command-and-control host) are labelled
    assert {int(r["port"]) for r in got if r["tags"]} == {oracle_netsimple.MalwareModel.BAD_NEWS_BEAR_PORT}


def _reference_schemas():
    reference_schema = "/root/reference/itsim/schemas/items.py"
    if not os.path.exists(reference_schema):
        return None
    scope = {}
    source = open(reference_schema).read().split("def get_itsim_object_types")[0].replace("import warlock", "")
    exec(compile(source, reference_schema, "exec"), scope)          # the schema dicts only (warlock is not installed)
    return scope


def test_bulk_tables_equal_the_per_record_documents(hostemu, tmp_path):
    """The vectorised converters (whole drains at once) give, record for record, the documents of the per-record ones,
    valid against the reference's three schemas (network_event, node, log), loadable with one executemany per table;
    and they do it at a rate a drain of millions of records can live with."""
    import time
    from itsim_b200 import netsimple, telemetry
    sim_uuid = "11111111-2222-3333-4444-555555555555"
    sim = netsimple.build(connection_records=60000, connection_window=8, trace_records=60000)
    sim.caps["net_event_records"] = 20000
    sim._lib = hostemu
    eng = sim.run_replicas(3, seed=0)
    eng.run(1200.0)
    trace = eng.trace(1)
    d = eng.drain(copy=True)
    rep = telemetry._replica_column(d["net_first"], len(d["net"]))
    assert len(rep) == len(d["net"]) and rep.max() == 2
    table = telemetry.network_event_table(d["net"], sim_uuid, replica_of=rep)
    docs = [json.loads(x) for x in table["json"]]
    slow = telemetry.network_event_documents(d["net"], sim_uuid=sim_uuid)
    assert len(docs) == len(slow) > 300
    for fast, ref in zip(docs, slow):
        for key in ("sim_uuid", "timestamp", "type", "network_event_type", "pid", "protocol", "src", "dst"):
            assert fast[key] == ref[key], key
    assert len(set(table["uuid"].tolist())) == len(docs)               # ids are unique
    nodes = telemetry.node_table(sim, sim_uuid, replicas=3)
    node_ids = set(nodes["uuid"].tolist())
    assert len(node_ids) == 3 * len(sim.nodes) and {x["uuid_node"] for x in docs} <= node_ids
    assert json.loads(nodes["json"][2])["node_label"] == "Workstation-1 #0"
    logs = telemetry.log_table(trace, sim_uuid, program_names={netsimple.PROG_MDNS: "mdns_daemon"})
    assert len(logs["json"]) == len(trace) and "mdns_daemon" in " ".join(logs["json"][:2000].tolist())
    schemas = _reference_schemas()
    if schemas is not None:
        import jsonschema
        for x in docs[:40]:
            jsonschema.validate(x, schemas["NETWORK_EVENT_SCHEMA"])
        for x in nodes["json"][:40]:
            jsonschema.validate(json.loads(x), schemas["NODE_SCHEMA"])
        for x in logs["json"][:40]:
            jsonschema.validate(json.loads(x), schemas["LOG_SCHEMA"])
    db = sqlite3.connect(str(tmp_path / "itsim.sqlite"))
    for name, tab in (("network_event", table), ("node", nodes), ("log", logs)):
        db.execute(f"CREATE TABLE {name} (uuid TEXT PRIMARY KEY, timestamp TEXT, sim_uuid TEXT, json TEXT)")
        db.executemany(f"INSERT INTO {name} VALUES (?, ?, ?, ?)", telemetry.table_rows(tab))
        assert db.execute(f"SELECT COUNT(*) FROM {name}").fetchone()[0] == len(tab["uuid"])
    # the connection lines of Socket.log_pair
    lines = telemetry.connection_jsonl(d["conn"])
    slow_conn = telemetry.connection_documents(d["conn"])
    assert len(lines) == len(slow_conn) > 3000
    for k in range(0, len(lines), 37):
        assert json.loads(lines[k]) == slow_conn[k]
    big = np.tile(d["conn"], 1 + 200000 // len(d["conn"]))[:200000]
    t0 = time.perf_counter()
    out = telemetry.connection_jsonl(big)
    rate = len(out) / (time.perf_counter() - t0)
    assert json.loads(out[-1]) == telemetry.connection_documents(big[-1:])[0]
    assert rate > 5e4, rate          # the per-record form manages ~1e5/s before json.dumps; this must not be slower
