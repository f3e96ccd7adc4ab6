"""Pins the Philox4x32-10 stream the oracle and the engine share (Random123 known-answer vectors)."""
import random

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
    """Only the bit source is swapped: variates are random.Random's own methods."""
    rng = PhiloxRandom(3, 4)
    assert type(rng).expovariate is random.Random.expovariate
    assert type(rng).normalvariate is random.Random.normalvariate
    a, _ = draw_words(3, 4, 0)
    assert rng.getrandbits(6) == a >> 26
    x = rng.choice(list(range(49)))
    assert 0 <= x < 49
    state = rng.getstate()
    v = rng.random()
    rng.setstate(state)
    assert rng.random() == v
