"""The engine's random stream restated (csrc/search.h: `Rng`): xoshiro256** seeded through splitmix64, 53-bit uniforms,
Marsaglia's polar normal, Marsaglia-Tsang gamma (shape < 1 boosted by U^(1/shape), as rand_distr does) and the
Dirichlet sample of `dirichlet.rs:48-70`.  TEST INFRASTRUCTURE ONLY.  The reference draws from `thread_rng()`; there is
nothing to pin against -- this exists so that the ORACLE search can consume the same numbers as the product and whole
self-play games can be compared move for move."""
from __future__ import annotations

import math

import numpy as np

M64 = (1 << 64) - 1


def _rotl(x: int, k: int) -> int:
    return ((x << k) | (x >> (64 - k))) & M64


class Rng:
    def __init__(self, seed: int = 1):
        self.reseed(seed)

    def reseed(self, seed: int) -> None:
        self.s = []
        seed &= M64
        for _ in range(4):
            seed = (seed + 0x9e3779b97f4a7c15) & M64
            z = seed
            z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & M64
            z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & M64
            self.s.append(z ^ (z >> 31))

    def next(self) -> int:
        s = self.s
        r = (_rotl((s[1] * 5) & M64, 7) * 9) & M64
        t = (s[1] << 17) & M64
        s[2] ^= s[0]
        s[3] ^= s[1]
        s[1] ^= s[2]
        s[0] ^= s[3]
        s[2] ^= t
        s[3] = _rotl(s[3], 45)
        return r

    def uniform(self) -> float:
        return float(self.next() >> 11) * (1.0 / 9007199254740992.0)

    def below(self, n: int) -> int:
        return int(self.uniform() * n)

    def normal(self) -> float:
        while True:
            u = 2.0 * self.uniform() - 1.0
            v = 2.0 * self.uniform() - 1.0
            q = u * u + v * v
            if 0.0 < q < 1.0:
                return u * math.sqrt(-2.0 * math.log(q) / q)

    def gamma(self, shape: float) -> float:
        if shape < 1.0:
            u = self.uniform()
            while u <= 0.0:
                u = self.uniform()
            return self.gamma(shape + 1.0) * math.pow(u, 1.0 / shape)
        d = shape - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * d)
        while True:
            x = self.normal()
            v = 1.0 + c * x
            if v <= 0.0:
                continue
            v = v * v * v
            u = self.uniform()
            if u < 1.0 - 0.0331 * x * x * x * x:
                return d * v
            if math.log(u) < 0.5 * x * x + d * (1.0 - v + math.log(v)):
                return d * v

    def dirichlet(self, x: np.ndarray, shape: float) -> np.ndarray:
        """eta[i] = g_i / sum g over the finite entries of x (dirichlet.rs:48-70)."""
        while True:
            g = [0.0] * 362
            total, count = 0.0, 0
            for i in range(362):
                if np.isfinite(x[i]):
                    g[i] = self.gamma(shape)
                    total += g[i]
                    count += 1
            if count == 0 or total > 2.2250738585072014e-308:
                break
        return np.array([np.float32(gi / total) if count else np.float32(0.0) for gi in g], np.float32)
