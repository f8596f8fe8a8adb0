"""ctypes front-end of the libdg_go restatement in `oracle/dg_oracle_go.cpp`.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU-baseline legs) -- never imported by the product.
Points are packed indices 19*y + x with (x, y) as in `Point::new(x, y)` (libdg_go/point.rs:26-34); 361 = pass.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import oracle as _nn

BLACK, WHITE = 1, 2
# order of symmetry::ALL (libdg_go/utils/symmetry.rs:121-130)
IDENTITY, FLIP_LR, FLIP_UD, TRANSPOSE, TRANSPOSE_ANTI, ROT90, ROT180, ROT270 = range(8)
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "go_fixtures.npz")
_ready = False


def lib() -> C.CDLL:
    global _ready
    L = _nn.lib()
    if not _ready:
        P, I, F, U64 = C.c_void_p, C.c_int, C.c_float, C.c_uint64
        sig = {
            "dgo_set_zobrist_table": (None, [P]), "dgo_reset_zobrist_table": (None, []),
            "dgo_board_new": (P, [F]), "dgo_board_clone": (P, [P]), "dgo_board_free": (None, [P]),
            "dgo_board_set_komi": (None, [P, F]), "dgo_board_komi": (F, [P]),
            "dgo_board_place": (None, [P, I, I]), "dgo_board_is_valid": (I, [P, I, I]),
            "dgo_board_is_valid_fast": (I, [P, I, I]), "dgo_board_is_ko": (I, [P, I, I]),
            "dgo_board_at": (I, [P, I]), "dgo_board_zobrist_hash": (U64, [P]), "dgo_board_to_move": (I, [P]),
            "dgo_board_count": (I, [P]), "dgo_board_get_n_liberty": (I, [P, I]),
            "dgo_board_get_n_liberty_if": (I, [P, I, I]), "dgo_board_is_ladder_capture": (I, [P, I, I]),
            "dgo_board_is_ladder_escape": (I, [P, I, I]), "dgo_board_stones": (None, [P, P]),
            "dgo_board_legal_mask": (None, [P, I, P]), "dgo_board_features_v1": (None, [P, I, I, P]),
            "dgo_board_is_symmetric": (I, [P, I]), "dgo_symmetry_apply": (I, [I, I]), "dgo_symmetry_inverse": (I, [I]),
            "dgo_ladder_nodes": (C.c_long, []), "dgo_f32_to_f16": (C.c_uint16, [F]),
            "dgo_replay": (I, [F, P, P, I, P, P, P]),
            "dgo_board_benson": (None, [P, I, P]), "dgo_board_territory": (None, [P, P]), "dgo_board_is_scorable": (I, [P]),
            "dgo_board_policy_candidates": (None, [P, I, I, P]), "dgo_board_is_simple_eye": (I, [P, I, I]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _ready = True
    return L


def use_reference_zobrist() -> None:
    """Loads the reference's zobrist constants (data fixture) so hashes compare with real_games.rs."""
    table = np.ascontiguousarray(np.load(_GOLDEN)["zobrist"], np.uint64)
    lib().dgo_set_zobrist_table(table.ctypes.data)


def use_default_zobrist() -> None:
    """Back to the splitmix64 constants the product also uses (hashes then compare product vs oracle)."""
    lib().dgo_reset_zobrist_table()


def idx(x: int, y: int) -> int:
    return 19 * y + x


class Board:
    """`dg_go::Board` (libdg_go/board.rs)."""

    def __init__(self, komi: float = 7.5, _handle=None):
        self._h = _handle if _handle is not None else lib().dgo_board_new(komi)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dgo_board_free(self._h)
            self._h = None

    def clone(self) -> "Board":
        return Board(_handle=lib().dgo_board_clone(self._h))

    def place(self, color: int, x: int, y: int) -> None:
        lib().dgo_board_place(self._h, color, idx(x, y))

    def place_index(self, color: int, index: int) -> None:
        lib().dgo_board_place(self._h, color, index)

    def is_valid(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dgo_board_is_valid(self._h, color, idx(x, y)))

    def is_valid_fast(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dgo_board_is_valid_fast(self._h, color, idx(x, y)))

    def at(self, x: int, y: int) -> int:
        return lib().dgo_board_at(self._h, idx(x, y))

    def zobrist_hash(self) -> int:
        return int(lib().dgo_board_zobrist_hash(self._h))

    def to_move(self) -> int:
        return lib().dgo_board_to_move(self._h)

    def count(self) -> int:
        return lib().dgo_board_count(self._h)

    def get_n_liberty(self, x: int, y: int) -> int:
        return lib().dgo_board_get_n_liberty(self._h, idx(x, y))

    def get_n_liberty_if(self, color: int, x: int, y: int) -> int:
        return lib().dgo_board_get_n_liberty_if(self._h, color, idx(x, y))

    def is_ladder_capture(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dgo_board_is_ladder_capture(self._h, color, idx(x, y)))

    def is_ladder_escape(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dgo_board_is_ladder_escape(self._h, color, idx(x, y)))

    def stones(self) -> np.ndarray:
        out = np.empty(361, np.uint8)
        lib().dgo_board_stones(self._h, out.ctypes.data)
        return out

    def legal_mask(self, color: int) -> np.ndarray:
        out = np.empty(361, np.uint8)
        lib().dgo_board_legal_mask(self._h, color, out.ctypes.data)
        return out

    def features(self, to_move: int, symmetry: int = IDENTITY) -> np.ndarray:
        """`features::V1::get_features::<HWC, f16>` -> [361, 32] fp16."""
        out = np.empty((361, 32), np.float16)
        lib().dgo_board_features_v1(self._h, to_move, symmetry, out.ctypes.data)
        return out

    def is_symmetric(self, transform: int) -> bool:
        return bool(lib().dgo_board_is_symmetric(self._h, transform))

    def benson(self, color: int) -> np.ndarray:
        """0 none / 1 unconditionally alive / 2 vital region per point (utils/benson.rs:152-175)."""
        out = np.empty(361, np.uint8)
        lib().dgo_board_benson(self._h, color, out.ctypes.data)
        return out

    def is_scorable(self) -> bool:
        return bool(lib().dgo_board_is_scorable(self._h))

    def territory(self) -> np.ndarray:
        """1 / 2 / 0 per point: whose territory the game record counts it as (score.rs:148-195)."""
        out = np.empty(361, np.uint8)
        lib().dgo_board_territory(self._h, out.ctypes.data)
        return out

    def result(self) -> str:
        """`get_winner_as_sgf` (game_result.rs:78-93)."""
        t = self.territory()
        black = np.float32((t == 1).sum())
        white = np.float32(np.float32((t == 2).sum()) + np.float32(lib_komi(self)))
        if black > white:
            return "B+%.1f" % float(np.float32(black - white))
        if white > black:
            return "W+%.1f" % float(np.float32(white - black))
        return "0"

    def policy_candidates(self, to_move: int, kind: int = 0) -> np.ndarray:
        out = np.empty(362, np.uint8)
        lib().dgo_board_policy_candidates(self._h, to_move, kind, out.ctypes.data)
        return out

    def is_simple_eye(self, color: int, x: int, y: int) -> bool:
        return bool(lib().dgo_board_is_simple_eye(self._h, color, idx(x, y)))


def lib_komi(board: "Board") -> float:
    return float(lib().dgo_board_komi(board._h))


def symmetry_apply(transform: int, index: int) -> int:
    return lib().dgo_symmetry_apply(transform, index)


def replay(colors: np.ndarray, moves: np.ndarray, komi: float = 7.5, features: bool = False, legal: bool = False,
           hashes: bool = False):
    """Replays one game; returns dict with per-ply 'features' [n,361,32] fp16 / 'legal' [n,361] / 'hash' [n]
    (position BEFORE each move, to_move = the move's colour).  Raises on an illegal move."""
    colors = np.ascontiguousarray(colors, np.uint8)
    moves = np.ascontiguousarray(moves, np.uint16)
    n = len(moves)
    out = {}
    f = np.empty((n, 361, 32), np.float16) if features else None
    l = np.empty((n, 361), np.uint8) if legal else None
    h = np.empty(n, np.uint64) if hashes else None
    rc = lib().dgo_replay(komi, colors.ctypes.data, moves.ctypes.data, n,
                          f.ctypes.data if features else None, l.ctypes.data if legal else None,
                          h.ctypes.data if hashes else None)
    if rc < 0:
        raise ValueError(f"illegal move at ply {-rc - 1}")
    if features:
        out["features"] = f
    if legal:
        out["legal"] = l
    if hashes:
        out["hash"] = h
    return out


def load_games():
    """The 99 fixture games as a list of (colors u8[n], moves u16[n], komi)."""
    z = np.load(_GOLDEN)
    off = z["games_offsets"]
    return [(z["games_colors"][off[i]:off[i + 1]], z["games_moves"][off[i]:off[i + 1]], float(z["games_komi"][i]))
            for i in range(len(off) - 1)]
