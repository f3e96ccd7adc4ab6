"""Bodies the front-end must refuse (tests/test_frontend.py), each with the text its CompileError has to contain."""
import inspect

from itsim.simulator import advance, now
from itsim.types import Protocol


def first_line(f):
    return inspect.getsourcelines(f)[1]


def five_handles(context):
    a = context.thread.clone(idle)
    b = context.thread.clone(idle)
    c = context.thread.clone(idle)
    d = context.thread.clone(idle)
    e = context.thread.clone(idle)
    for t in (a, b, c, d, e):
        t.join()


def idle(context):
    advance(1)


def clock_in_arithmetic(context):
    deadline = now() + 5
    advance(deadline)


def loop_over_packets(context):
    with context.node.bind(Protocol.UDP, 99) as sock:
        packet = sock.recv()
        for _ in range(packet.byte_size):
            advance(1)


def stale_packet(context):
    with context.node.bind(Protocol.UDP, 99) as sock:
        one = sock.recv()
        two = sock.recv()
        sock.send(one.source, 8)


def packet_after_blocking(context):
    with context.node.bind(Protocol.UDP, 99) as sock:
        packet = sock.recv()
        advance(1)
        sock.send(packet.source, 8)


def rich_payload(context):
    with context.node.bind(Protocol.UDP) as sock:
        sock.send(("10.0.0.1", 5), 8, {"content": "x", "blob": [1, 2, 3]})


def new_exception(context):
    advance(1)
    raise RuntimeError("boom")


tally = []


def host_state_steers(context):
    with context.node.bind(Protocol.UDP, 99) as sock:
        while True:
            sock.recv()
            tally.append(1)
            if len(tally) > 3:          # the reference sees the live list; a compiled body would decide this once
                break


def generator_body(context):
    yield 1


REFUSED = [
    (five_handles, "registers"),
    (clock_in_arithmetic, "now()"),
    (loop_over_packets, "mixes simulation values"),
    (stale_packet, "no longer the last one received"),
    (packet_after_blocking, "do not survive a blocking call"),
    (rich_payload, "payload keys"),
    (new_exception, "raising a new exception"),
    (host_state_steers, "cannot steer the simulation"),
    (generator_body, "not supported"),
]
