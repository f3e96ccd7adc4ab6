"""Unit constants, quirks included, exactly as ``itsim/units.py:2-20`` computes them (US = MS*1e-6 = 1e-9 s,
NS = US*1e-9; bandwidth constants are bytes per second although named bits)."""
S = 1.0
MIN = 60 * S
H = 60 * MIN
D = 24 * H
W = 7 * D
MS = S * 1.0e-3
US = MS * 1.0e-6
NS = US * 1.0e-9

KbPS = 1024 / 8
MbPS = 1024 * KbPS
GbPS = 1024 * MbPS

B = 1
KB = 1024 * B
MB = 1024 * KB
GB = 1024 * MB
