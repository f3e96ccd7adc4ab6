"""
Declarative variate generators with the reference's names and chaining (``greensim.random`` as used by
``itsim/random.py:14-29``, ``itsim/network/link.py:40-41`` and ``examples/network_simple/network_simple.py:101-103,
128,171,249-250``).  The reference's generators are opaque Python iterators; to be compiled for the device they
have to say what they are, so each factory returns a :class:`Dist` descriptor that the model compiler lowers to an
``itsb_dist`` row.  All draws of a replica come from its one Philox stream (key = (seed, replica)) in event order.
"""
from typing import Optional

CONSTANT, UNIFORM, EXPO, NORMAL, CHOICE = 0, 1, 2, 3, 4
DF_LINEAR, DF_LOWER, DF_UPPER, DF_INT = 1, 2, 4, 8


class Dist:
    """One generator chain: base variate -> linear -> bounded -> project_int (the only shapes ITsim builds)."""

    def __init__(self, kind: int, p0: float, p1: float = 0.0) -> None:
        self.kind, self.p0, self.p1 = kind, float(p0), float(p1)
        self.flags = 0
        self.slope, self.shift, self.lower, self.upper = 1.0, 0.0, 0.0, 0.0

    def _copy(self) -> "Dist":
        d = Dist(self.kind, self.p0, self.p1)
        d.flags, d.slope, d.shift, d.lower, d.upper = self.flags, self.slope, self.shift, self.lower, self.upper
        return d

    def key(self) -> tuple:
        return (self.kind, self.p0, self.p1, self.flags, self.slope, self.shift, self.lower, self.upper)


def constant(value: float) -> Dist:
    return Dist(CONSTANT, value)


def uniform(lower: float, upper: float) -> Dist:
    return Dist(UNIFORM, lower, upper)


def expo(mean: float) -> Dist:
    """``rng.expovariate(1.0 / mean)``: the rate is computed here exactly as greensim computes it."""
    return Dist(EXPO, 1.0 / mean)


def normal(mean: float, std_dev: float) -> Dist:
    return Dist(NORMAL, mean, std_dev)


def distribution_index(n: int) -> Dist:
    """``rng.choice(seq)`` reduced to the index it picks (``greensim.random.distribution``)."""
    return Dist(CHOICE, n)


def linear(gen: Dist, slope: float, shift: float) -> Dist:
    if gen.flags:
        raise ValueError("linear() must be applied first (the order itsim.random.num_bytes uses)")
    d = gen._copy()
    d.flags |= DF_LINEAR
    d.slope, d.shift = float(slope), float(shift)
    return d


def bounded(gen: Dist, lower: Optional[float] = None, upper: Optional[float] = None) -> Dist:
    if gen.flags & DF_INT:
        raise ValueError("bounded() after project_int()")
    d = gen._copy()
    # nested clamps compose exactly (clamping is monotone): the Internet latency of overrides.py:443-456 is
    # bounded(bounded(expo(1 s), lower=20 ms), lower=0)
    if lower is not None:
        d.lower = max(d.lower, float(lower)) if d.flags & DF_LOWER else float(lower)
        d.flags |= DF_LOWER
    if upper is not None:
        d.upper = min(d.upper, float(upper)) if d.flags & DF_UPPER else float(upper)
        d.flags |= DF_UPPER
    if (d.flags & DF_LOWER) and (d.flags & DF_UPPER) and d.lower > d.upper:
        raise ValueError("empty clamp interval")
    return d


def project_int(gen: Dist) -> Dist:
    d = gen._copy()
    d.flags |= DF_INT
    return d


def num_bytes(gen: Dist, header: int = 0, upper: Optional[float] = None) -> Dist:
    """``itsim/random.py:14-29``: ``project_int(bounded(linear(gen, 1.0, header), 0.0, upper))``."""
    return project_int(bounded(linear(gen, 1.0, header), 0.0, upper))
