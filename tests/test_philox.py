"""Pins the Philox4x32-10 stream the oracle and the engine share (Random123 known-answer vectors), and the one
logarithm both sides compute their variates with."""
import inspect
import math
import random
import struct

import numpy as np
import pytest

from philox import MASK, PhiloxRandom, draw_words, philox4x32_10

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((MASK,) * 4, (MASK, MASK), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_random123_known_answers():
    for ctr, key, expected in KAT:
        assert philox4x32_10(ctr, key) == expected


def test_stream_layout():
    rng = PhiloxRandom(seed=(5 << 32) | 9, replica=(1 << 32) | 77)
    for d in range(6):
        block = d >> 1
        words = philox4x32_10((block, 0, 5, 1), (9, 77))
        a, b = words[2 * (d & 1)], words[2 * (d & 1) + 1]
        assert draw_words((5 << 32) | 9, (1 << 32) | 77, d) == (a, b)
        assert rng.random() == ((a >> 5) * 67108864 + (b >> 6)) / 9007199254740992.0
    assert rng.draws == 6


def test_transforms_are_cpythons():
    """The bit source is swapped, and the logarithm inside expovariate / normalvariate (see oracle/philox.py): every
    other variate is random.Random's own method, and the two overridden ones are CPython's formulas -- checked here
    against the stock methods on the same uniforms with math.log put back."""
    import philox
    rng = PhiloxRandom(3, 4)
    assert type(rng).uniform is random.Random.uniform
    assert type(rng).choice is random.Random.choice
    assert "NV_MAGICCONST*(u1-0.5)/u2" in inspect.getsource(random.Random.normalvariate).replace(" ", "")
    assert philox._NV_MAGICCONST == random.NV_MAGICCONST

    class Stock(PhiloxRandom):
        expovariate = random.Random.expovariate
        normalvariate = random.Random.normalvariate

    saved = philox.log_f64
    philox.log_f64 = math.log
    try:
        ours, stock = PhiloxRandom(8, 1), Stock(8, 1)
        for _ in range(20000):
            assert ours.normalvariate(100.0, 30.0) == stock.normalvariate(100.0, 30.0)
            assert ours.expovariate(1.0 / 192.0) == stock.expovariate(1.0 / 192.0)
        assert ours.draws == stock.draws
    finally:
        philox.log_f64 = saved
    a, _ = draw_words(3, 4, 0)
    assert rng.getrandbits(6) == a >> 26
    x = rng.choice(list(range(49)))
    assert 0 <= x < 49
    state = rng.getstate()
    v = rng.random()
    rng.setstate(state)
    assert rng.random() == v


def _log_inputs(n=200000):
    """What the variates feed the logarithm (1 - random() in (0, 1]) plus the corners of the algorithm."""
    rng = PhiloxRandom(77, 5)
    x = [1.0 - rng.random() for _ in range(n)]
    gen = random.Random(3)
    x += [gen.random() * 10.0 ** gen.uniform(-300, 300) for _ in range(n // 4)]
    x += [1.0, 2.0, 0.5, 1.0 - 2.0 ** -53, 2.0 ** -53, 1.0 + 2.0 ** -30, 1.0 - 2.0 ** -30, 1.0 + 2.0 ** -21,
          5e-324, 1e-310, 2.2250738585072014e-308, 1.7976931348623157e308, 0.7071067811865475, 0.7071067811865476,
          1.4142135623730951, 1.414213562373095, float("inf")]
    return np.array(x, dtype=np.float64)


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


def test_log_is_within_one_ulp_of_libm():
    from philox import log_f64
    x = _log_inputs(40000)
    ours = np.array([log_f64(float(v)) for v in x])
    libm = np.array([math.log(float(v)) for v in x])
    fin = np.isfinite(libm)
    assert np.array_equal(ours[~fin], libm[~fin])
    ulps = np.abs(_bits(ours[fin]).astype(np.int64) - _bits(libm[fin]).astype(np.int64))
    assert ulps.max() <= 1
    assert log_f64(0.0) == float("-inf") and math.isnan(log_f64(-1.0)) and log_f64(1.0) == 0.0


def test_engine_log_equals_oracle_log_bit_for_bit_host_build(hostemu):
    """The engine source (compiled for the CPU) and the oracle's Python compute the same logarithm, every bit."""
    from philox import log_f64
    x = _log_inputs(100000)
    y = np.empty_like(x)
    assert hostemu.itsb_eval_log(None, x.ctypes.data, y.ctypes.data, len(x)) == 0
    ref = np.array([log_f64(float(v)) for v in x])
    assert np.array_equal(_bits(y), _bits(ref))


@pytest.mark.gpu
def test_engine_log_equals_oracle_log_bit_for_bit_on_the_device():
    """Same check through the C ABI on the B200: the kernel's log_f64 (no FMA: -fmad=false) against the oracle's."""
    import ctypes
    from itsim_b200 import _lib
    from philox import log_f64
    lib = _lib.load()
    handle = ctypes.c_void_p()
    assert lib.itsb_create(0, ctypes.byref(handle)) == 0
    x = _log_inputs(300000)
    y = np.empty_like(x)
    assert lib.itsb_eval_log(handle, x.ctypes.data, y.ctypes.data, len(x)) == 0
    lib.itsb_destroy(handle)
    ref = np.array([log_f64(float(v)) for v in x])
    assert np.array_equal(_bits(y), _bits(ref))
