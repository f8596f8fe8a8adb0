"""Host-side mirror of the reference's `dg_mcts` crate surface for the self-play hot path, over the C ABI of
`include/dg_mcts.h` (product code: `csrc/search.h`, `csrc/search_task.h`, `csrc/search_api.cpp`).

* `predict(predictor, options, board, color, starting_tree)`  -- `dg_mcts::predict` (src/libdg_mcts/lib.rs:145-200)
* `Tree`                                                       -- `tree::Node` (forward / disqualify / counts)
* `self_play(predictor, ...)`                                  -- `dg_mcts::self_play` (self_play.rs:423-500)
* `EnginePredictor(network)`                                   -- `predictors::nn::NnPredictor` (predictors/nn.rs:84-107)
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import numpy as np

from . import go, nn

PREDICT_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p)
PREDICT_RAW_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p)
PREDICT_PRIOR_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


class _SearchOptions(C.Structure):
    _fields_ = [("search", C.c_int32), ("deterministic", C.c_int32), ("num_rollout", C.c_int32),
                ("probes_per_round", C.c_int32), ("dirichlet_noise", C.c_float), ("temperature", C.c_float),
                ("seed", C.c_uint64), ("noise", C.c_void_p), ("leaf_symmetries", C.c_void_p),
                ("n_leaf_symmetries", C.c_int32), ("choose_at", C.c_double), ("cache", C.c_void_p), ("device_ladders", C.c_int32)]


class _SelfPlayConfig(C.Structure):
    _fields_ = [("num_games", C.c_int32), ("num_parallel", C.c_int32), ("num_rollout", C.c_int32),
                ("probes_per_round", C.c_int32), ("max_plies", C.c_int32), ("num_threads", C.c_int32),
                ("ex_it", C.c_int32), ("num_ex_it_rollout", C.c_int32), ("dirichlet_noise", C.c_float),
                ("temperature", C.c_float), ("seed", C.c_uint64), ("max_seconds", C.c_double),
                ("cache_capacity", C.c_int32), ("num_groups", C.c_int32), ("cache_shared", C.c_int32)]


class _SelfPlayStats(C.Structure):
    _fields_ = [("games_finished", C.c_int64), ("moves", C.c_int64), ("evals", C.c_int64), ("rounds", C.c_int64),
                ("searches", C.c_int64), ("seconds", C.c_double), ("eval_seconds", C.c_double),
                ("mean_batch", C.c_double), ("digest", C.c_uint64), ("cache_hits", C.c_int64)]


_P, _I = C.c_void_p, C.c_int32
# every symbol include/dg_mcts.h declares
ABI = {
    "dg_engine_predict": (_I, [_P, _P, _I, _P, _P]),
    "dg_random_predict": (_I, [_P, _P, _I, _P, _P]),
    "dg_peaked_predict": (_I, [_P, _P, _I, _P, _P]),
    "dg_engine_predict_raw": (_I, [_P, _P, _I, _P, _P, _P]),
    "dg_engine_predict_prior": (_I, [_P, _P, _I, _P, _P, _P, _P]),
    "dg_mcts_predict_prior": (_I, [PREDICT_PRIOR_FN, _P, C.POINTER(_SearchOptions), _P, _P, _I, C.POINTER(C.c_float),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "dg_selfplay_run_prior": (_I, [PREDICT_PRIOR_FN, _P, C.POINTER(_SelfPlayConfig), C.POINTER(_SelfPlayStats), _P, C.c_int64]),
    "dg_mcts_predict_raw": (_I, [PREDICT_RAW_FN, _P, C.POINTER(_SearchOptions), _P, _P, _I, C.POINTER(C.c_float),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "dg_selfplay_run_raw": (_I, [PREDICT_RAW_FN, _P, C.POINTER(_SelfPlayConfig), C.POINTER(_SelfPlayStats), _P, C.c_int64]),
    "dg_mcts_predict": (_I, [PREDICT_FN, _P, C.POINTER(_SearchOptions), _P, _P, _I, C.POINTER(C.c_float),
                             C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "dg_cache_new": (_P, [_I]), "dg_cache_new_shared": (_P, [_I, _I]), "dg_cache_free": (None, [_P]),
    "dg_cache_stats": (None, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "dg_tree_free": (None, [_P]), "dg_tree_forward": (_P, [_P, _I]), "dg_tree_disqualify": (None, [_P, _I]),
    "dg_tree_total_count": (_I, [_P]), "dg_tree_to_move": (_I, [_P]), "dg_tree_initial_value": (C.c_float, [_P]),
    "dg_tree_children": (None, [_P, _P, _P, _P]), "dg_tree_num_nodes": (C.c_int64, [_P]),
    "dg_selfplay_run": (_I, [PREDICT_FN, _P, C.POINTER(_SelfPlayConfig), C.POINTER(_SelfPlayStats), _P, C.c_int64]),
    "dg_selfplay_run_engine": (_I, [C.POINTER(C.c_void_p), _I, C.c_uint32, C.POINTER(_SelfPlayConfig), C.POINTER(_SelfPlayStats), _P,
                                    C.c_int64]),
}
SELFPLAY_DEVICE_PRIORS = 0x1
SELFPLAY_DEVICE_LADDERS = 0x2
SELFPLAY_AUTO_PRIORS = 0x4
_ready = False


def lib() -> C.CDLL:
    global _ready
    L = go.lib()
    if not _ready:
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _ready = True
    return L


def python_predictor(fn: Callable[[np.ndarray], tuple]):
    """Wraps `fn(features [n,361,32] fp16) -> (value [n] fp16, policy [n,362] fp16)` as a dg_predict_fn (tests, stubs)."""
    def call(_ctx, positions, n, value, policy):
        try:
            buf = (C.c_uint8 * (n * nn.PACKED_DTYPE.itemsize)).from_address(positions)
            packed = np.frombuffer(buf, dtype=nn.PACKED_DTYPE)
            v, p = fn(go.unpack_features(packed))
            C.memmove(value, np.ascontiguousarray(v, np.float16).ctypes.data, 2 * n)
            C.memmove(policy, np.ascontiguousarray(p, np.float16).ctypes.data, 2 * 362 * n)
            return 0
        except Exception:   # noqa: BLE001 -- must not unwind through C
            import traceback
            traceback.print_exc()
            return -2
    return PREDICT_FN(call)


class EnginePredictor:
    """The product predictor: leaves go to the B200 engine (`dg_engine_forward_packed`)."""

    def __init__(self, network: "nn.Network"):
        self.network = network
        self.fn = C.cast(lib().dg_engine_predict, PREDICT_FN)
        self.ctx = network._handle


class EngineRawPredictor:
    """The engine fed with raw positions: feature planes and legal moves are computed on the device."""
    raw = True

    def __init__(self, network: "nn.Network"):
        self.network = network
        self.fn = C.cast(lib().dg_engine_predict_raw, PREDICT_RAW_FN)
        self.ctx = network._handle


class EnginePriorPredictor:
    """The engine fed with raw positions, returning ready-to-insert priors as well (planes, legal moves, candidate masks,
    inverse symmetry and renormalisation all on the device)."""
    raw = "prior"

    def __init__(self, network: "nn.Network"):
        self.network = network
        self.fn = C.cast(lib().dg_engine_predict_prior, PREDICT_PRIOR_FN)
        self.ctx = network._handle


class EngineQueue:
    """The product path of self-play: one or several engines (one per device) driven through their leaf-batch queues by
    `dg_selfplay_run_engine` -- no blocking predictor call.  `device_priors`: build the leaves' priors on the device
    (None = the driver decides batch by batch from how busy its worker threads are, DG_SELFPLAY_AUTO_PRIORS); `device_ladders`: the
    ladder planes are read on the device as well (None = with a single host thread per engine)."""
    engines = True

    def __init__(self, networks, device_priors=None, device_ladders=None):
        self.networks = list(networks) if isinstance(networks, (list, tuple)) else [networks]
        self.device_priors = device_priors
        self.device_ladders = device_ladders


class RandomPredictor:
    """`predictors::RandomPredictor` as a deterministic function of the position (host only, no device)."""

    def __init__(self):
        self.fn = C.cast(lib().dg_random_predict, PREDICT_FN)
        self.ctx = None


class PeakedPredictor:
    """A peaked stand-in for a trained network (host only): searches go deep instead of wide."""

    def __init__(self, sharpness: float = 12.0):
        self._sharp = C.c_float(sharpness)
        self.fn = C.cast(lib().dg_peaked_predict, PREDICT_FN)
        self.ctx = C.cast(C.pointer(self._sharp), C.c_void_p)


class Cache:
    """Transposition table of evaluations (`NnPredictor`'s `LruCache`, predictors/nn.rs:29-82)."""

    def __init__(self, capacity: int = 200_000, stripes: int = 0):
        """stripes = 0: one LRU list behind one lock (the reference's table); n > 0: n lock stripes (shared by many searches)."""
        self._h = lib().dg_cache_new_shared(capacity, stripes) if stripes > 0 else lib().dg_cache_new(capacity)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dg_cache_free(self._h)
            self._h = None

    def stats(self):
        hits, misses, size = C.c_int64(), C.c_int64(), C.c_int64()
        lib().dg_cache_stats(self._h, C.byref(hits), C.byref(misses), C.byref(size))
        return {"hits": hits.value, "misses": misses.value, "size": size.value}


class Tree:
    """`tree::Node` owned by the caller."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dg_tree_free(self._h)
            self._h = None

    def release(self):
        h, self._h = self._h, None
        return h

    def forward(self, index: int) -> Optional["Tree"]:
        """`Node::forward` (tree.rs:1198-1225): consumes this tree."""
        h = lib().dg_tree_forward(self.release(), index)
        return Tree(h) if h else None

    def disqualify(self, index: int) -> None:
        lib().dg_tree_disqualify(self._h, index)

    @property
    def total_count(self) -> int:
        return lib().dg_tree_total_count(self._h)

    @property
    def to_move(self) -> int:
        return lib().dg_tree_to_move(self._h)

    @property
    def initial_value(self) -> float:
        return float(lib().dg_tree_initial_value(self._h))

    def num_nodes(self) -> int:
        return int(lib().dg_tree_num_nodes(self._h))

    def children(self):
        count = np.empty(362, np.int32)
        value = np.empty(362, np.float32)
        prior = np.empty(362, np.float32)
        lib().dg_tree_children(self._h, count.ctypes.data, value.ctypes.data, prior.ctypes.data)
        return count, value, prior


def _fn_ctx(predictor):
    if hasattr(predictor, "fn") and hasattr(predictor, "ctx"):      # a native predictor: function pointer + context
        return predictor.fn, predictor.ctx
    return predictor, None


def predict(predictor, board: "go.Board", color: int, *, search: int = go.STANDARD_SEARCH, deterministic: bool = False,
            num_rollout: int = 800, probes_per_round: int = 1, starting_tree: Optional[Tree] = None, seed: int = 1,
            noise: Optional[np.ndarray] = None, dirichlet_noise: float = 0.25, temperature: float = 0.8,
            leaf_symmetries=None, choose_at: float = -1.0, cache: Optional["Cache"] = None, device_ladders: bool = False):
    """`dg_mcts::predict`.  Returns (value, index, Tree, evals)."""
    fn, ctx = _fn_ctx(predictor)
    opt = _SearchOptions(search, int(deterministic), num_rollout, probes_per_round, dirichlet_noise, temperature, seed,
                         None, None, 0, choose_at, cache._h if cache is not None else None, int(device_ladders))
    keep = []
    if noise is not None:
        eta = np.ascontiguousarray(noise, np.float32)
        keep.append(eta)
        opt.noise = eta.ctypes.data
    if leaf_symmetries is not None:
        ls = np.ascontiguousarray(leaf_symmetries, np.uint8)
        keep.append(ls)
        opt.leaf_symmetries = ls.ctypes.data
        opt.n_leaf_symmetries = len(ls)
    value, index, tree, evals = C.c_float(), C.c_int32(), C.c_void_p(), C.c_int64()
    kind = getattr(predictor, "raw", False)
    call = lib().dg_mcts_predict_prior if kind == "prior" else lib().dg_mcts_predict_raw if kind else lib().dg_mcts_predict
    rc = call(fn, ctx, C.byref(opt), starting_tree.release() if starting_tree is not None else None,
              board._h, color, C.byref(value), C.byref(index), C.byref(tree), C.byref(evals))
    if rc:
        raise nn.Error(rc, "predictor failed")
    return value.value, index.value, Tree(tree.value), evals.value


def self_play(predictor, *, num_games: int, num_parallel: int = 32, num_rollout: int = 800, probes_per_round: int = 8,
              max_plies: int = 722, num_threads: int = 0, ex_it: bool = False, num_ex_it_rollout: int = 800,
              dirichlet_noise: float = 0.25, temperature: float = 0.8, seed: int = 1, max_seconds: float = 0.0,
              cache_capacity: int = 0, num_groups: int = 0, sgf_capacity: int = 1 << 24, cache_shared: int = 0):
    """`dg_mcts::self_play`: returns (stats dict, list of SGF records)."""
    cfg = _SelfPlayConfig(num_games, num_parallel, num_rollout, probes_per_round, max_plies, num_threads, int(ex_it),
                          num_ex_it_rollout, dirichlet_noise, temperature, seed, max_seconds, cache_capacity, num_groups, cache_shared)
    stats = _SelfPlayStats()
    buf = C.create_string_buffer(sgf_capacity)
    if isinstance(predictor, EngineQueue):
        import os
        nets = predictor.networks
        handles = (C.c_void_p * len(nets))(*[n._handle for n in nets])
        # what else moves to the device is a question of which side is scarce, and that changes during a run (a leaf of the
        # middle game costs the host twice what a leaf of the opening costs): by default the driver decides batch by batch
        # (DG_SELFPLAY_AUTO_PRIORS), starting on the device when host threads are few
        threads = num_threads if num_threads > 0 else (os.cpu_count() or 1)
        priors, ladders = predictor.device_priors, predictor.device_ladders
        flags = 0
        if priors is None:
            flags |= SELFPLAY_AUTO_PRIORS | (SELFPLAY_DEVICE_PRIORS if threads < 8 * len(nets) else 0)
        elif priors:
            flags |= SELFPLAY_DEVICE_PRIORS
        if ladders is None:
            ladders = threads < 2 * len(nets)        # (measured: even with 2 host threads per engine the host reader wins)
        if ladders:
            flags |= SELFPLAY_DEVICE_LADDERS
        rc = lib().dg_selfplay_run_engine(handles, len(nets), flags, C.byref(cfg), C.byref(stats), buf, sgf_capacity)
    else:
        fn, ctx = _fn_ctx(predictor)
        kind = getattr(predictor, "raw", False)
        run = lib().dg_selfplay_run_prior if kind == "prior" else lib().dg_selfplay_run_raw if kind else lib().dg_selfplay_run
        rc = run(fn, ctx, C.byref(cfg), C.byref(stats), buf, sgf_capacity)
    if rc:
        raise nn.Error(rc, "self-play failed")
    out = {name: getattr(stats, name) for name, _ in _SelfPlayStats._fields_}
    return out, [g for g in buf.value.decode().split("\n") if g]
