"""Host logic: IR assembler, generator descriptors, model tables, validation at the C ABI."""
import ctypes as C

import numpy as np
import pytest


def test_assembler_labels_and_encoding():
    from itsim_b200 import ir
    a = ir.Asm()
    a.program("p", 9)
    a.label("top")
    a.recv(handler="h")
    a.jeqi(ir.PKT_TAG, 5, "top")
    a.advance(3, handler="h")
    a.label("h")
    a.exit()
    code, entries = a.assemble()
    assert [int(code[0]) & 0xFF, int(code[1]) & 0xFF] == [ir.OP_CLOSE, ir.OP_EXIT]   # reserved: unhandled Interrupt
    base = 2
    assert list(entries[0]) == [base, 9]
    op, imm = int(code[base]) & 0xFF, int(code[base]) >> 32
    assert op == ir.OP_RECV and imm >> 16 == base + 3
    op, imm = int(code[base + 1]) & 0xFF, int(code[base + 1]) >> 32
    assert op == ir.OP_JEQI and imm & 0xFFFF == base and imm >> 16 == 5
    op, imm = int(code[base + 2]) & 0xFF, int(code[base + 2]) >> 32
    assert op == ir.OP_ADVANCE and imm & 0xFFFF == 3 and imm >> 16 == base + 3
    with pytest.raises(KeyError):
        b = ir.Asm()
        b.program("q", 1)
        b.jmp("nowhere")
        b.assemble()


def test_num_bytes_chain_matches_reference_shape():
    """itsim/random.py:29: project_int(bounded(linear(gen, 1.0, header), 0.0, upper))."""
    from itsim_b200 import random as R
    d = R.num_bytes(R.expo(192.0), header=68, upper=576)
    assert d.kind == R.EXPO and d.p0 == 1.0 / 192.0
    assert d.flags == R.DF_LINEAR | R.DF_LOWER | R.DF_UPPER | R.DF_INT
    assert (d.slope, d.shift, d.lower, d.upper) == (1.0, 68.0, 0.0, 576.0)
    d = R.num_bytes(R.normal(100.0, 30.0), header=240)
    assert d.flags == R.DF_LINEAR | R.DF_LOWER | R.DF_INT


def test_units_quirks():
    """itsim/units.py:7-14 verbatim: US is 1e-9 s and GbPS is bytes per second."""
    from itsim_b200 import units as U
    assert U.US == U.MS * 1.0e-6 and U.US == pytest.approx(1e-9)
    assert U.GbPS == 134217728.0
    assert U.H == 3600.0


def test_network_simple_tables():
    from itsim_b200 import netsimple, _lib as L
    sim = netsimple.build()
    c = sim.compile()
    assert len(c.nodes) == 51 and len(c.init) == 148          # 148 resident processes
    assert c.nodes["addr"][0] == (192 << 24) | (168 << 16) | (4 << 8) | 1     # Router = .1, DHCP = .2, WS-k = .(k+2)
    assert c.nodes["addr"][50] == (192 << 24) | (168 << 16) | (4 << 8) | 51
    assert c.nets["bcast"][0] == (192 << 24) | (168 << 16) | (4 << 8) | 255
    mdns_entry, llmnr_entry = sim.asm.entry("mdns"), sim.asm.entry("llmnr")
    daemons = [int(e) for e in c.init["entry"] if int(e) in (mdns_entry, llmnr_entry)]
    assert daemons.count(mdns_entry) == 45 and daemons.count(llmnr_entry) == 4    # sic: n <= num_mdns
    lat = c.dists[int(c.nets["lat_dist"][0])]
    assert lat["kind"] == 3 and lat["p0"] == 5e-3 and lat["flags"] & 2 and lat["lower"] == 0.0
    bw = c.dists[int(c.nets["bw_dist"][0])]
    assert bw["p0"] == 134217728.0 and bw["lower"] == 0.125


def test_bad_models_are_rejected(hostemu):
    from itsim_b200 import netsimple, _lib as L
    from itsim_b200.engine import Engine
    sim = netsimple.build()
    c = sim.compile()
    c.code = c.code.copy()
    c.code[5] = 0xFF
    with pytest.raises(L.EngineError) as err:
        Engine(c, lib=hostemu)
    assert err.value.code == -2
    c = sim.compile()
    c.nets = c.nets.copy()
    c.nets["n_nodes"][0] = 999
    with pytest.raises(L.EngineError):
        Engine(c, lib=hostemu)


def test_network_link_errors():
    from itsim_b200.model import AddressInUse, Endpoint, InvalidAddress, Network, Simulator
    from itsim_b200.random import constant
    sim = Simulator()
    net = Network(sim, "10.0.0.0/30", constant(0.0), constant(1.0))
    Endpoint("a", net, "10.0.0.1")
    with pytest.raises(AddressInUse):
        Endpoint("b", net, "10.0.0.1")
    with pytest.raises(InvalidAddress):
        Endpoint("c", net, "10.0.1.1")
