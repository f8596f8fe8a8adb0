"""Shared helpers of the search tests: deterministic stub predictors over fp16 feature tensors."""
import zlib

import numpy as np


def hash_predictor(scale: float = 2.0, value_scale: float = 0.5):
    """Pseudo-random but deterministic function of the feature tensor (so it also depends on the symmetry)."""
    def fn(feats: np.ndarray):
        n = feats.shape[0]
        v = np.empty(n, np.float16)
        p = np.empty((n, 362), np.float16)
        for i in range(n):
            rng = np.random.default_rng(zlib.crc32(np.ascontiguousarray(feats[i]).tobytes()))
            logits = rng.normal(size=362) * scale
            e = np.exp(logits - logits.max())
            p[i] = (e / e.sum()).astype(np.float16)
            v[i] = np.float16(np.tanh(rng.normal() * value_scale))
        return v, p
    return fn


def fake_predictor(point: int, value: float):
    """`FakePredictor` (src/libdg_mcts/predictors/fake.rs:23-54)."""
    def fn(feats: np.ndarray):
        n = feats.shape[0]
        p = np.zeros((n, 362), np.float16)
        p[:, point] = 1.0
        return np.full(n, value, np.float16), p
    return fn


def nan_predictor():
    """`NanPredictor` (predictors/nan.rs:33-44): value 0, policy -inf everywhere."""
    def fn(feats: np.ndarray):
        n = feats.shape[0]
        return np.zeros(n, np.float16), np.full((n, 362), -np.inf, np.float16)
    return fn


def uniform_predictor(value: float = 0.0):
    def fn(feats: np.ndarray):
        n = feats.shape[0]
        return np.full(n, value, np.float16), np.full((n, 362), 1.0 / 362, np.float16)
    return fn


def dirichlet_sample(seed: int, prior: np.ndarray, shape: float = 0.03) -> np.ndarray:
    rng = np.random.default_rng(seed)
    g = np.where(np.isfinite(prior), rng.gamma(shape, size=len(prior)), 0.0)
    return (g / g.sum()).astype(np.float32)
