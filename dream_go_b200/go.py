"""Host-side mirror of the reference's `dg_go` crate surface for the self-play hot path, over the C ABI of
`include/dg_go.h` (product code: `csrc/go_board.h`, `csrc/go_api.cpp`).

Same names and argument meaning as the reference: `Board::{new, place, is_valid, at, to_move, count,
zobrist_hash}` (src/libdg_go/board.rs), `features::V1::get_features` (utils/features.rs:154-250),
`symmetry::{Transform, is_symmetric}` (utils/symmetry.rs), the ladder reader (utils/ladder.rs) and the
prior construction of `pool/policy_helper.rs`.  Points are `(x, y)` as in `Point::new(x, y)`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import nn

BLACK, WHITE, PASS = 1, 2, 361
STANDARD_SEARCH, SCORING_SEARCH = 0, 1
IDENTITY, FLIP_LR, FLIP_UD, TRANSPOSE, TRANSPOSE_ANTI, ROT90, ROT180, ROT270 = range(8)   # symmetry::ALL

_P, _I, _F = C.c_void_p, C.c_int32, C.c_float
# every symbol include/dg_go.h declares
ABI = {
    "dg_board_new": (_P, [_F]), "dg_board_clone": (_P, [_P]), "dg_board_copy": (None, [_P, _P]),
    "dg_board_free": (None, [_P]), "dg_board_set_komi": (None, [_P, _F]), "dg_board_komi": (_F, [_P]),
    "dg_board_count": (_I, [_P]), "dg_board_zobrist_hash": (C.c_uint64, [_P]), "dg_board_to_move": (_I, [_P]),
    "dg_board_at": (_I, [_P, _I]), "dg_board_is_valid": (_I, [_P, _I, _I]), "dg_board_place": (None, [_P, _I, _I]),
    "dg_board_get_n_liberty": (_I, [_P, _I]), "dg_board_get_n_liberty_if": (_I, [_P, _I, _I]),
    "dg_board_is_ladder_capture": (_I, [_P, _I, _I]), "dg_board_is_ladder_escape": (_I, [_P, _I, _I]),
    "dg_board_is_symmetric": (_I, [_P, _I]), "dg_board_legal_moves": (None, [_P, _I, _P]),
    "dg_symmetry_apply": (_I, [_I, _I]), "dg_symmetry_inverse": (_I, [_I]),
    "dg_board_features_packed": (None, [_P, _I, _I, _P, _P]), "dg_board_raw_position": (None, [_P, _I, _I, _P]), "dg_board_features_f16": (None, [_P, _I, _I, _P]),
    "dg_go_extract_batch": (None, [_P, _P, _P, _I, _P, _P, _I]),
    "dg_go_replay": (_I, [_F, _P, _P, _I, _P, _P, _P]),
    "dg_board_prior": (None, [_P, _I, _I, _P, _P, _I, _F, _P]),
    "dg_board_is_scorable": (_I, [_P]), "dg_board_territory": (None, [_P, _P]), "dg_board_benson": (None, [_P, _I, _P]),
    "dg_board_policy_candidates": (None, [_P, _I, _I, _P, _P]),
    "dg_go_set_zobrist": (None, [_P]),
}
_ready = False


def lib() -> C.CDLL:
    global _ready
    L = nn.lib()
    if not _ready:
        for name, (res, args) in ABI.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _ready = True
    return L


def set_zobrist(table) -> None:
    """`zobrist::TABLE` ([3][420] u64, src/libdg_go/zobrist.rs:18) for every hash of the process; None = built-in table."""
    if table is None:
        lib().dg_go_set_zobrist(None)
    else:
        t = np.ascontiguousarray(table, np.uint64)
        assert t.shape == (3, 420)
        lib().dg_go_set_zobrist(t.ctypes.data)


def idx(x: int, y: int) -> int:
    return 19 * y + x


class Board:
    """`dg_go::Board`."""

    def __init__(self, komi: float = 7.5, _handle=None):
        self._h = _handle if _handle is not None else lib().dg_board_new(komi)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dg_board_free(self._h)
            self._h = None

    def clone(self) -> "Board":
        return Board(_handle=lib().dg_board_clone(self._h))

    def copy_from(self, other: "Board") -> None:
        """`*self = other.clone()` without an allocation (the searches copy a board per probe)."""
        lib().dg_board_copy(self._h, other._h)

    def set_komi(self, komi: float) -> None:
        """`Board::set_komi` (board.rs:84-86)."""
        lib().dg_board_set_komi(self._h, komi)

    def place(self, color: int, x: int, y: int) -> None:
        lib().dg_board_place(self._h, color, idx(x, y))

    def place_index(self, color: int, index: int) -> None:
        lib().dg_board_place(self._h, color, index)

    def is_valid(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dg_board_is_valid(self._h, color, idx(x, y)))

    def at(self, x: int, y: int) -> int:
        return lib().dg_board_at(self._h, idx(x, y))

    def zobrist_hash(self) -> int:
        return int(lib().dg_board_zobrist_hash(self._h))

    def to_move(self) -> int:
        return lib().dg_board_to_move(self._h)

    def count(self) -> int:
        return lib().dg_board_count(self._h)

    def komi(self) -> float:
        return float(lib().dg_board_komi(self._h))

    def get_n_liberty(self, x: int, y: int) -> int:
        return lib().dg_board_get_n_liberty(self._h, idx(x, y))

    def get_n_liberty_if(self, color: int, x: int, y: int) -> int:
        return lib().dg_board_get_n_liberty_if(self._h, color, idx(x, y))

    def is_ladder_capture(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dg_board_is_ladder_capture(self._h, color, idx(x, y)))

    def is_ladder_escape(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dg_board_is_ladder_escape(self._h, color, idx(x, y)))

    def is_symmetric(self, transform: int) -> bool:
        return bool(lib().dg_board_is_symmetric(self._h, transform))

    def stones(self) -> np.ndarray:
        return np.array([lib().dg_board_at(self._h, i) for i in range(361)], np.uint8)

    def legal_moves(self, color: int) -> np.ndarray:
        out = np.empty(361, np.uint8)
        lib().dg_board_legal_moves(self._h, color, out.ctypes.data)
        return out

    def features_packed(self, to_move: int, symmetry: int = IDENTITY, legal: bool = False):
        """V1 features as one `dg_packed_position` (what `dg_engine_forward_packed` / the leaf queue take)."""
        out = np.zeros(1, nn.PACKED_DTYPE)
        lg = np.empty(361, np.uint8) if legal else None
        lib().dg_board_features_packed(self._h, to_move, symmetry, out.ctypes.data, lg.ctypes.data if legal else None)
        return (out, lg) if legal else out

    def raw_position(self, to_move: int, symmetry: int = IDENTITY, search: int = STANDARD_SEARCH) -> np.ndarray:
        """One `dg_raw_position` (what `dg_engine_forward_raw` takes): the device derives planes and legal moves; `search`
        (bits 4.. of the symmetry byte) only matters to `dg_engine_forward_raw_prior`."""
        out = np.zeros(1, nn.RAW_DTYPE)
        lib().dg_board_raw_position(self._h, to_move, symmetry | (search << 4), out.ctypes.data)
        return out

    def features(self, to_move: int, symmetry: int = IDENTITY) -> np.ndarray:
        """`features::V1::get_features::<HWC, f16>` -> [361, 32] fp16."""
        out = np.empty((361, 32), np.float16)
        lib().dg_board_features_f16(self._h, to_move, symmetry, out.ctypes.data)
        return out

    def is_scorable(self) -> bool:
        """`Score::is_scorable` (utils/score.rs:97-110)."""
        return bool(lib().dg_board_is_scorable(self._h))

    def benson(self, color: int) -> np.ndarray:
        out = np.empty(361, np.uint8)
        lib().dg_board_benson(self._h, color, out.ctypes.data)
        return out

    def territory(self) -> np.ndarray:
        """Per point 1 / 2 / 0: whose territory the game record counts it as (`get_stone_status`, score.rs:148-195)."""
        out = np.empty(361, np.uint8)
        lib().dg_board_territory(self._h, out.ctypes.data)
        return out

    def policy_candidates(self, to_move: int, search: int = STANDARD_SEARCH) -> np.ndarray:
        out = np.empty(362, np.uint8)
        lib().dg_board_policy_candidates(self._h, to_move, search, None, out.ctypes.data)
        return out

    def prior(self, to_move: int, policy: np.ndarray, symmetry: int = IDENTITY, sum_to: float = 1.0, legal=None,
              search: int = STANDARD_SEARCH) -> np.ndarray:
        """create_initial_policy + add_valid_candidates + normalize_policy (pool/worker_thread.rs:88-93)."""
        policy = np.ascontiguousarray(policy, np.float16)
        assert policy.shape == (362,)
        out = np.empty(368, np.float32)
        lg = None if legal is None else np.ascontiguousarray(legal, np.uint8)
        lib().dg_board_prior(self._h, to_move, search, None if lg is None else lg.ctypes.data, policy.ctypes.data, symmetry,
                             sum_to, out.ctypes.data)
        return out


def symmetry_apply(transform: int, index: int) -> int:
    return lib().dg_symmetry_apply(transform, index)


def unpack_features(packed: np.ndarray) -> np.ndarray:
    """dg_packed_position[n] -> [n, 361, 32] fp16 (host restatement of the GPU pack kernel, for tests)."""
    planes = packed["planes"].astype(np.uint32)                       # [n, 361]
    bits = ((planes[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.float16)
    k = packed["k_bits"].astype(np.uint16).view(np.float16)
    bits[..., 0] *= k[:, None]
    bits[..., 1] *= k[:, None]
    return bits


def replay(colors, moves, komi: float = 7.5, features: bool = False, legal: bool = False, hashes: bool = False):
    colors = np.ascontiguousarray(colors, np.uint8)
    moves = np.ascontiguousarray(moves, np.uint16)
    n = len(moves)
    f = np.zeros(n, nn.PACKED_DTYPE) if features else None
    l = np.empty((n, 361), np.uint8) if legal else None
    h = np.empty(n, np.uint64) if hashes else None
    rc = lib().dg_go_replay(komi, colors.ctypes.data, moves.ctypes.data, n, f.ctypes.data if features else None,
                            l.ctypes.data if legal else None, h.ctypes.data if hashes else None)
    if rc < 0:
        raise ValueError(f"illegal move at ply {-rc - 1}")
    out = {}
    if features:
        out["features"] = f
    if legal:
        out["legal"] = l
    if hashes:
        out["hash"] = h
    return out


class PreparedBatch:
    """Arguments of `dg_go_extract_batch` marshalled once (benchmarks time `run()` alone, not the ctypes set-up)."""

    def __init__(self, boards, to_move, legal: bool = True, threads: int = 0):
        self.n = len(boards)
        self.boards = boards
        self.handles = (C.c_void_p * self.n)(*[b._h for b in boards])
        self.tm = np.ascontiguousarray(to_move, np.uint8)
        self.out = np.zeros(self.n, nn.PACKED_DTYPE)
        self.legal = np.empty((self.n, 361), np.uint8) if legal else None
        self.threads = threads
        self._fn = lib().dg_go_extract_batch

    def run(self):
        self._fn(self.handles, self.tm.ctypes.data, None, self.n, self.out.ctypes.data,
                 self.legal.ctypes.data if self.legal is not None else None, self.threads)


def extract_batch(boards, to_move, symmetry=None, legal: bool = False, threads: int = 0):
    """BASELINE.json configs[0]: features + legal moves for a batch of boards on the host cores."""
    n = len(boards)
    handles = (C.c_void_p * n)(*[b._h for b in boards])
    tm = np.ascontiguousarray(to_move, np.uint8)
    sy = None if symmetry is None else np.ascontiguousarray(symmetry, np.uint8)
    out = np.zeros(n, nn.PACKED_DTYPE)
    lg = np.empty((n, 361), np.uint8) if legal else None
    lib().dg_go_extract_batch(handles, tm.ctypes.data, None if sy is None else sy.ctypes.data, n, out.ctypes.data,
                              lg.ctypes.data if legal else None, threads)
    return (out, lg) if legal else out
