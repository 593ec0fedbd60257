"""java.util.Random restated from the public JDK specification (oracle; test infra).

The reference draws corrupt entities with ``new Random().nextInt(numOfentities)``
(DataFactory.java:119,128) and the per-relation corrupt side with
``Math.random() > 0.5`` (NTN.java:146).  Both are unseeded there; the host owns
the seeds here so that indices are bit-exact against what a seeded JVM produces.

Algorithm (JDK javadoc of java.util.Random):
    seed0   = (seed ^ 0x5DEECE66D) & (2^48 - 1)
    next(b) : s = (s * 0x5DEECE66D + 0xB) & (2^48 - 1); return (int)(s >>> (48 - b))
"""
from __future__ import annotations

_MULT = 0x5DEECE66D
_ADD = 0xB
_MASK = (1 << 48) - 1


def _to_int32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


class JavaRandom:
    def __init__(self, seed: int):
        self.s = (seed ^ _MULT) & _MASK

    def next(self, bits: int) -> int:
        self.s = (self.s * _MULT + _ADD) & _MASK
        return _to_int32(self.s >> (48 - bits))

    def next_int(self, bound: int | None = None) -> int:
        if bound is None:
            return self.next(32)
        if bound <= 0:
            raise ValueError("bound must be positive")
        r = self.next(31)
        m = bound - 1
        if (bound & m) == 0:  # power of two
            return _to_int32((bound * r) >> 31)
        u = r
        while True:
            r = u % bound
            # Java: while (u - r + m < 0) in int32 arithmetic (overflow check)
            if _to_int32(u - r + m) >= 0:
                return r
            u = self.next(31)

    def next_double(self) -> float:
        return ((self.next(26) << 27) + self.next(27)) * (1.0 / (1 << 53))
