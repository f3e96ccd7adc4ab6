"""
Process bodies written the way an ITsim user writes them -- plain Python functions over ``context`` -- for the
front-end tests.  They exercise the same primitives, in the same order, as the scenarios whose golden traces
(tests/golden/machine_*.npz) were produced by the reference's integration-test bodies on the oracle, so a trace
compiled from these functions must equal those goldens event for event.  Imports are the reference's module paths:
the tests load this file with ``itsim_b200.compat`` installed.
"""
from greensim.random import constant, uniform

from itsim.machine.endpoint import Endpoint
from itsim.network.link import Link
from itsim.simulator import advance, now
from itsim.types import Protocol, Timeout, as_address
from itsim.machine.process_management.thread import ThreadKilled
from itsim.units import MS, MbPS, S

# ---- ping / pong over one link (golden: machine_packet_transfer_*.npz) ------------------------------------------------
PONG_PORT = 9887
seen = set()
clock_at_timeout = []


def pinger(context):
    with context.node.bind(Protocol.UDP) as sock:
        sock.send(("10.11.12.20", PONG_PORT), 4, {"content": "ping"})
        try:
            answer = sock.recv(1 * S)
        except Timeout:
            raise_failure("no answer within a second")
        assert answer.byte_size == 8
        assert answer.payload["content"] == "pong"
        try:
            sock.recv(5 * S)
            raise_failure("nobody sends a second packet")
        except Timeout:
            clock_at_timeout.append(now())
            assert now() > 5.0
        finally:
            seen.add("pinger")


def ponger(context):
    with context.node.bind(Protocol.UDP, PONG_PORT) as sock:
        request = sock.recv()
        assert request.payload["content"] == "ping"
        sock.send(request.source, 8, {"content": "pong"})
        seen.add("ponger")


def raise_failure(why):
    raise AssertionError(why)


def build_ping_pong(sim):
    """Every builder returns its nodes in creation order (the order the traces number them in)."""
    link = Link("10.11.12.0/24", latency=uniform(100 * MS, 200 * MS), bandwidth=constant(100 * MbPS))
    a = Endpoint().connected_to_static(link, "10.11.12.10")
    a.run_proc_in(sim, 0.1, pinger)
    b = Endpoint().connected_to_static(link, "10.11.12.20")
    b.run_proc(sim, ponger)
    return [a, b]


# ---- twenty endpoints broadcasting (golden: machine_broadcast_*.npz) -------------------------------------------------
class Chatterbox(Endpoint):
    PORT = 10000

    def __init__(self, sim):
        super().__init__()
        self.heard_from = set()
        self.gap = uniform(3.0 * S, 6.0 * S)
        self.run_proc(sim, self.listen)
        self.run_proc(sim, self.shout)

    def listen(self, context):
        with context.node.bind(Protocol.UDP, self.PORT) as sock:
            while True:
                packet = sock.recv()
                self.heard_from.add(packet.source.hostname_as_address())

    def shout(self, context):
        with context.node.bind(Protocol.UDP) as sock:
            while True:
                advance(next(self.gap))
                for interface in self.interfaces():
                    if interface.address.is_loopback:
                        continue
                    sock.send((interface.cidr.broadcast_address, self.PORT), 128)


def build_chatter(sim, hosts=range(100, 120)):
    link = Link("10.11.0.0/16", uniform(100 * MS, 200 * MS), constant(100 * MbPS))
    return [Chatterbox(sim).connected_to_static(link, as_address(n, link.cidr)) for n in hosts]


def build_threads(sim, log):
    return [Endpoint().with_proc_in(sim, 0, root, log)]


def build_family(sim, graves):
    return [Endpoint().with_proc(sim, ancestors, graves)]


# ---- threads of one process (golden: machine_thread_lifecycle.npz) ---------------------------------------------------
def observer(context, log, watched):
    watched.wait()
    log.append("observer")


def leaf(context, pid, log):
    assert context.process.pid == pid
    advance(20)
    log.append("leaf")


def first(context, pid, log):
    assert context.process.pid == pid
    t = context.thread.clone(leaf, pid, log)
    t.join()
    log.append("first")


def second(context, pid, log):
    assert context.process.pid == pid
    advance(5)
    log.append("second")


def sleeper(context, pid, log):
    try:
        advance(1000)
        log.append("sleeper woke up by itself")
    except ThreadKilled:
        raise
    finally:
        log.append("sleeper")


def root(context, log):
    context.process.fork_exec(observer, log, context.process)
    t1 = context.thread.clone(first, context.process.pid, log)
    t2 = context.thread.clone(second, context.process.pid, log)
    t3 = context.thread.clone(sleeper, context.process.pid, log)
    t1.join()
    t2.join()
    try:
        t3.join(100)
        log.append("the sleeper cannot be done yet")
    except Timeout:
        pass
    log.append("root")
    context.process.kill()


# ---- a family of processes (golden: machine_process_lifecycle.npz) ---------------------------------------------------
class Being:
    def __init__(self, name, born, dies):
        self.name, self.born, self.dies = name, born, dies

    def life(self, context, graves, family):
        try:
            assert context.process is family[self.name]
            self.live(context, family)
        finally:
            graves.add(self.name)

    def live(self, context, family):
        advance(self.dies - self.born)


class Killer(Being):
    def live(self, context, family):
        advance(130 - now())
        family["victim"].kill()
        advance(self.dies - now())


def ancestors(context, graves):
    family = {}
    beings = [Killer("killer", 0, 730), Being("victim", 0, 1000), Being("third", 130, 130 + 912),
              Being("crowd", 0, 100000)]
    for b in beings:
        family[b.name] = context.process.fork_exec_in(b.born, b.life, graves, family)
    for name, year in (("victim", 130), ("killer", 730), ("third", 1042)):
        family[name].wait()
        assert abs(now() - year) < 1e-9
    try:
        family["crowd"].wait(2000)
        graves.add("the crowd cannot be gone")
    except Timeout:
        pass
    graves.add("ancestors")


# ---- config C4's traffic: many endpoints, one program each for servers and clients (golden: machine_large_lan_*.npz) --
LAN_PORT = 10000
lan_gap = uniform(3.0, 6.0)


def lan_server(context):
    with context.node.bind(Protocol.UDP, LAN_PORT) as sock:
        while True:
            packet = sock.recv()
            if packet.payload.get("content") == "request":
                sock.send(packet.source, 32, {"content": "reply"})


def lan_client(context, peer):
    with context.node.bind(Protocol.UDP) as sock:
        while True:
            advance(next(lan_gap))
            for interface in context.node.interfaces():
                if not interface.address.is_loopback:
                    sock.send((interface.cidr.broadcast_address, LAN_PORT), 128, {"content": "chatter"})
            sock.send((peer, LAN_PORT), 64, {"content": "request"})


def build_lan(sim, subnets, per_subnet, per_process, make_router):
    from greensim.random import normal
    from itsim.network.route import Relay
    from itsim.units import GbPS
    backbone = Link("10.30.0.0/24", normal(5 * MS, 1 * MS), constant(1 * GbPS))
    access = [Link(f"10.20.{i}.0/25", normal(5 * MS, 1 * MS), constant(1 * GbPS)) for i in range(subnets)]
    nodes = []
    for i in range(subnets):
        router = make_router()
        nodes.append(router)
        router.connected_to_static(access[i], 1)
        router.connected_to_static(backbone, i + 1, [Relay(as_address(j + 1, backbone.cidr), f"10.20.{j}.0/25")
                                                    for j in range(subnets) if j != i])
    for i in range(subnets):
        for h in range(per_subnet):
            ep = Endpoint().connected_to_static(access[i], h + 2, [Relay(as_address(1, access[i].cidr))])
            nodes.append(ep)
            ep.run_proc(sim, lan_server)
            ep.run_proc(sim, lan_client, per_process(as_address(h + 2, access[(i + 1) % subnets].cidr)))
    return nodes


# ---- a request / response protocol with fields in the payload (golden: frontend_echo_*.npz, made by running THIS file
# ---- on the reference's itsim: oracle/machine_scenarios.frontend_scenario) ----------------------------------------------
ECHO_PORT = 4000
echoed = []
unanswered = []


def asker(context, first_id):
    with context.node.bind(Protocol.UDP) as sock:
        for k in range(3):
            advance(next(ask_gap))
            sock.send(("10.5.0.2", ECHO_PORT), 40, {"content": "question", "id": first_id + k, "weight": 5 * k})
            try:
                reply = sock.recv(0.25 * S)
            except Timeout:
                unanswered.append(first_id + k)
                continue
            assert reply.payload["content"] == "answer"
            if reply.payload["id"] != first_id + k:
                unanswered.append(-1)
            echoed.append(reply.payload["weight"])


def answerer(context):
    with context.node.bind(Protocol.UDP, ECHO_PORT) as sock:
        while True:
            q = sock.recv()
            if q.payload["content"] != "question":
                continue
            if q.payload["weight"] >= 10:
                continue                      # the third question of every asker goes unanswered
            sock.send(q.source, 24, {"content": "answer", "id": q.payload["id"], "weight": q.payload["weight"]})


ask_gap = uniform(0.5 * S, 1.5 * S)


def build_echo(sim):
    link = Link("10.5.0.0/24", latency=uniform(10 * MS, 60 * MS), bandwidth=constant(10 * MbPS))
    server = Endpoint().connected_to_static(link, "10.5.0.2")
    server.run_proc(sim, answerer)
    nodes = [server]
    for n in range(3):
        ep = Endpoint().connected_to_static(link, as_address(10 + n, link.cidr))
        ep.run_proc_in(sim, 0.05 * n, asker, 100 * (n + 1))
        nodes.append(ep)
    return nodes


# ---- a networking daemon (Node.networking_daemon, node.py:350-392) answering on two ports (golden: frontend_daemon_*.npz)
served = []


def build_daemon(sim):
    link = Link("10.6.0.0/24", latency=uniform(20 * MS, 80 * MS), bandwidth=constant(10 * MbPS))
    server = Endpoint().connected_to_static(link, "10.6.0.2")

    @server.networking_daemon(sim, Protocol.UDP, 7, 9)
    def echo(context, packet, socket):
        size = packet.byte_size              # what is needed after the advance() is copied out of the packet first
        addr = packet.source.hostname_as_address()
        port = packet.source.port
        served.append(size)
        if packet.payload["content"] == "loud":
            advance(0.5 * S)                 # requests overlap: every packet is served by a thread of its own
        socket.send((addr, port), size, {"content": "echo"})

    nodes = [server]
    for n in range(2):
        ep = Endpoint().connected_to_static(link, as_address(20 + n, link.cidr))
        ep.run_proc_in(sim, 0.1 * (n + 1), caller, 7 + 2 * n, 32 * (n + 1))
        nodes.append(ep)
    return nodes


def caller(context, port, size):
    with context.node.bind(Protocol.UDP) as sock:
        for content in ("loud", "quiet", "loud"):
            sock.send(("10.6.0.2", port), size, {"content": content})
            advance(0.2 * S)
        for _ in range(3):
            answer = sock.recv(3 * S)
            assert answer.byte_size == size
