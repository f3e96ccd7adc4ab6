"""
ORACLE / TEST INFRASTRUCTURE ONLY.  ``greensim.random`` variate generators, as consumed by the reference:
``itsim/random.py:4,29`` (linear, bounded, project_int), ``itsim/network/link.py:4,40-41,92`` (constant, bounded),
``itsim/network/service/dhcp/client.py:4,47-49`` and ``server.py:5,73-75`` (normal), the examples (expo, uniform,
distribution) and the reference tests (uniform, constant).

Every generator is an infinite iterator.  ``rng=None`` means "the default source", which upstream is the
process-global Mersenne Twister; here the default is *resolved at each draw* through :func:`default_rng` so the
oracle can inject one Philox-backed stream per replica (SURVEY.md section 8-c "canonicalisation") even into
generators built at import time.  Variate transforms are CPython's ``random.Random`` methods, unmodified.
"""
import random as _random
from typing import Iterator, Optional, Sequence, TypeVar

T = TypeVar("T")

VarRandom = Iterator
RandomOpt = Optional[_random.Random]

_default = [None]


def set_default_rng(rng: RandomOpt) -> None:
    """Installs the source used by generators built with ``rng=None`` (None restores the global ``random``)."""
    _default[0] = rng


def default_rng():
    return _default[0] if _default[0] is not None else _random


def _src(rng: RandomOpt):
    return rng if rng is not None else default_rng()


def constant(value: T) -> Iterator[T]:
    while True:
        yield value


def linear(gen: Iterator[float], slope: float, shift: float) -> Iterator[float]:
    for x in gen:
        yield slope * x + shift


def bounded(gen: Iterator[float], lower: Optional[float] = None, upper: Optional[float] = None) -> Iterator[float]:
    for x in gen:
        if lower is not None and x < lower:
            x = lower
        if upper is not None and x > upper:
            x = upper
        yield x


def project_int(gen: Iterator[float]) -> Iterator[int]:
    for x in gen:
        yield int(x)


def uniform(lower: float, upper: float, rng: RandomOpt = None) -> Iterator[float]:
    while True:
        yield _src(rng).uniform(lower, upper)


def expo(mean: float, rng: RandomOpt = None) -> Iterator[float]:
    while True:
        yield _src(rng).expovariate(1.0 / mean)


def normal(mean: float, std_dev: float, rng: RandomOpt = None) -> Iterator[float]:
    while True:
        yield _src(rng).normalvariate(mean, std_dev)


def distribution(seq: Sequence[T], rng: RandomOpt = None) -> Iterator[T]:
    """Uniform choice among ``seq``, evaluated lazily at each draw (the reference fills the list after creating
    the generator: ``examples/network_simple/network_simple.py:369-375``)."""
    while True:
        yield _src(rng).choice(seq)
