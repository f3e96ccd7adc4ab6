"""
ORACLE / TEST INFRASTRUCTURE ONLY.

Counter-based random stream shared (by definition, not by code) with the CUDA engine
(``itsim_b200/csrc/itsb_philox.cuh``): Philox4x32-10 (Salmon et al., SC'11; constants as in Random123 and
``/usr/local/cuda/include/curand_philox4x32_x.h``), pinned by the Random123 known-answer vectors in
``tests/test_philox.py``.

Stream layout (the "RNG contract"):

* key     = (seed mod 2**32, replica mod 2**32)
* counter = (block mod 2**32, block >> 32, seed >> 32, replica >> 32),  block = draw >> 1
* draw ``d`` consumes the 64 bits ``(a, b) = words[2*(d&1)], words[2*(d&1)+1]`` of its block
* ``random()``      = ((a >> 5) * 2**26 + (b >> 6)) / 2**53       -- CPython's 53-bit construction
* ``getrandbits(k)``, k <= 32 = a >> (32 - k)                      -- one draw; wider requests chain draws, low word first

The reference draws every variate from CPython's ``random.Random`` methods (``uniform``, ``expovariate``,
``normalvariate``, ``choice``) through greensim's generators; :class:`PhiloxRandom` swaps the bit source and ONE
function: the natural logarithm inside ``expovariate`` and ``normalvariate``.  CPython calls the platform's libm
``log`` there, whose last bit is not specified (glibc and CUDA both promise <= 1 ulp, not the same ulp), and a
1-ulp difference flips ``int(min(max(x + 68, 0), 576))`` or the Kinderman-Monahan acceptance test once in ~1e13
draws, after which a replica's packet sizes and draw counts diverge for good.  So both sides compute the logarithm
with one algorithm, written with IEEE +, -, *, / only (no FMA, no table): :func:`log_f64` below is the same sequence
of operations as ``log_f64`` in ``itsim_b200/csrc/itsb_core.cuh`` (FreeBSD msun / fdlibm ``e_log.c``: argument
reduction to [sqrt(2)/2, sqrt(2)), degree-14 minimax polynomial in s = f / (2 + f), error < 1 ulp), and the two
variates are CPython's own formulas (Lib/random.py ``expovariate``, ``normalvariate``) with that logarithm.
``uniform``, ``choice`` and everything else run unmodified.  One stream per replica, consumed in event-execution
order.
"""
import random as _random
import struct as _struct
from typing import Tuple

_LN2_HI = float.fromhex("0x1.62e42fee00000p-1")    # 3fe62e42 fee00000
_LN2_LO = float.fromhex("0x1.a39ef35793c76p-33")   # 3dea39ef 35793c76
_LG1 = float.fromhex("0x1.5555555555593p-1")       # 3FE55555 55555593
_LG2 = float.fromhex("0x1.999999997fa04p-2")       # 3FD99999 9997FA04
_LG3 = float.fromhex("0x1.2492494229359p-2")       # 3FD24924 94229359
_LG4 = float.fromhex("0x1.c71c51d8e78afp-3")       # 3FCC71C5 1D8E78AF
_LG5 = float.fromhex("0x1.7466496cb03dep-3")       # 3FC74664 96CB03DE
_LG6 = float.fromhex("0x1.39a09d078c69fp-3")       # 3FC39A09 D078C69F
_LG7 = float.fromhex("0x1.2f112df3e5244p-3")       # 3FC2F112 DF3E5244
_TWO54 = 18014398509481984.0
_NV_MAGICCONST = 1.7155277699214135                # 4 * exp(-0.5) / sqrt(2.0), Lib/random.py


def log_f64(x: float) -> float:
    """Natural logarithm, bit-identical to ``log_f64`` of the CUDA engine (same operations in the same order)."""
    bits = _struct.unpack("<q", _struct.pack("<d", x))[0]
    hx = bits >> 32                       # signed high word
    lx = bits & 0xFFFFFFFF
    k = 0
    if hx < 0x00100000:                   # x < 2**-1022
        if ((hx & 0x7FFFFFFF) | lx) == 0:
            return float("-inf")
        if hx < 0:
            return float("nan")
        k -= 54
        x *= _TWO54
        bits = _struct.unpack("<q", _struct.pack("<d", x))[0]
        hx = bits >> 32
        lx = bits & 0xFFFFFFFF
    if hx >= 0x7FF00000:
        return x + x
    k += (hx >> 20) - 1023
    hx &= 0x000FFFFF
    i = (hx + 0x95F64) & 0x100000
    x = _struct.unpack("<d", _struct.pack("<q", ((hx | (i ^ 0x3FF00000)) << 32) | lx))[0]   # x or x / 2 in [sqrt(2)/2, sqrt(2))
    k += i >> 20
    f = x - 1.0
    dk = float(k)
    if (0x000FFFFF & (2 + hx)) < 3:       # |f| < 2**-20
        if f == 0.0:
            if k == 0:
                return 0.0
            return dk * _LN2_HI + dk * _LN2_LO
        r = f * f * (0.5 - 0.33333333333333333 * f)
        if k == 0:
            return f - r
        return dk * _LN2_HI - ((r - dk * _LN2_LO) - f)
    s = f / (2.0 + f)
    z = s * s
    i = hx - 0x6147A
    w = z * z
    j = 0x6B851 - hx
    t1 = w * (_LG2 + w * (_LG4 + w * _LG6))
    t2 = z * (_LG1 + w * (_LG3 + w * (_LG5 + w * _LG7)))
    i |= j
    r = t2 + t1
    if i > 0:
        hfsq = 0.5 * f * f
        if k == 0:
            return f - (hfsq - s * (hfsq + r))
        return dk * _LN2_HI - ((hfsq - (s * (hfsq + r) + dk * _LN2_LO)) - f)
    if k == 0:
        return f - s * (f - r)
    return dk * _LN2_HI - ((s * (f - r) - dk * _LN2_LO) - f)

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr: Tuple[int, int, int, int], key: Tuple[int, int]) -> Tuple[int, int, int, int]:
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for rnd in range(10):
        if rnd:
            k0 = (k0 + W0) & MASK
            k1 = (k1 + W1) & MASK
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
    return c0, c1, c2, c3


def draw_words(seed: int, replica: int, draw: int) -> Tuple[int, int]:
    """The 64 bits (a, b) consumed by draw number ``draw`` of stream (seed, replica)."""
    block = draw >> 1
    words = philox4x32_10((block & MASK, (block >> 32) & MASK, (seed >> 32) & MASK, (replica >> 32) & MASK),
                          (seed & MASK, replica & MASK))
    half = draw & 1
    return words[2 * half], words[2 * half + 1]


class PhiloxRandom(_random.Random):
    """``random.Random`` whose bits come from the (seed, replica) Philox stream.  ``draws`` = draws consumed."""

    def __init__(self, seed: int = 0, replica: int = 0) -> None:
        self._seed = int(seed)
        self._replica = int(replica)
        self.draws = 0
        self._block = -1
        self._words = (0, 0, 0, 0)
        super().__init__(0)

    # -- bit source ------------------------------------------------------------------------------------
    def _next(self) -> Tuple[int, int]:
        d = self.draws
        self.draws = d + 1
        block = d >> 1
        if block != self._block:
            self._block = block
            self._words = philox4x32_10(
                (block & MASK, (block >> 32) & MASK, (self._seed >> 32) & MASK, (self._replica >> 32) & MASK),
                (self._seed & MASK, self._replica & MASK))
        half = d & 1
        return self._words[2 * half], self._words[2 * half + 1]

    def random(self) -> float:
        a, b = self._next()
        return ((a >> 5) * 67108864 + (b >> 6)) / 9007199254740992.0

    def getrandbits(self, k: int) -> int:
        if k < 0:
            raise ValueError("number of bits must be non-negative")
        if k == 0:
            return 0
        if k <= 32:
            return self._next()[0] >> (32 - k)
        out, shift = 0, 0
        while k > 0:
            word = self._next()[0]
            if k < 32:
                word >>= 32 - k
            out |= word << shift
            shift += 32
            k -= 32
        return out

    # -- the two variates that take a logarithm: CPython's formulas with the shared log_f64 ----------------
    def expovariate(self, lambd: float = 1.0) -> float:
        """Lib/random.py ``expovariate``: ``-log(1.0 - random()) / lambd``."""
        return -log_f64(1.0 - self.random()) / lambd

    def normalvariate(self, mu: float = 0.0, sigma: float = 1.0) -> float:
        """Lib/random.py ``normalvariate`` (Kinderman-Monahan ratio of uniforms)."""
        while True:
            u1 = self.random()
            u2 = 1.0 - self.random()
            z = _NV_MAGICCONST * (u1 - 0.5) / u2
            zz = z * z / 4.0
            if zz <= -log_f64(u2):
                break
        return mu + z * sigma

    # -- state -----------------------------------------------------------------------------------------
    def seed(self, *args, **kwargs) -> None:  # the stream is defined by (seed, replica); reseeding rewinds it.
        self.draws = 0
        self._block = -1

    def getstate(self):
        return (self._seed, self._replica, self.draws)

    def setstate(self, state) -> None:
        self._seed, self._replica, self.draws = state
        self._block = -1
