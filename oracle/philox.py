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
``normalvariate``, ``choice``) through greensim's generators; :class:`PhiloxRandom` only swaps the bit source, so
those transforms run unmodified.  One stream per replica, consumed in event-execution order.
"""
import random as _random
from typing import Tuple

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

    # -- state -----------------------------------------------------------------------------------------
    def seed(self, *args, **kwargs) -> None:  # the stream is defined by (seed, replica); reseeding rewinds it.
        self.draws = 0
        self._block = -1

    def getstate(self):
        return (self._seed, self._replica, self.draws)

    def setstate(self, state) -> None:
        self._seed, self._replica, self.draws = state
        self._block = -1
