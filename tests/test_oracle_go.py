"""Pins the libdg_go restatement (oracle/dg_oracle_go.cpp) against every known-answer test the reference
holds for the rules / ladder / symmetry / feature path.  CPU only."""
import numpy as np
import pytest

from oracle import go
from oracle.go import BLACK, WHITE, Board


@pytest.fixture(scope="module", autouse=True)
def _zobrist():
    go.use_reference_zobrist()


# ---- libdg_go/board.rs:281-388 ----------------------------------------------------------------------------

def test_capture():
    b = Board()
    b.place(BLACK, 9, 9)
    for x, y in [(8, 9), (10, 9), (9, 8), (9, 10)]:
        b.place(WHITE, x, y)
    assert b.at(9, 9) == 0


def test_capture_group():
    b = Board()
    for x, y in [(0, 1), (1, 0), (0, 0), (1, 1)]:
        b.place(BLACK, x, y)
    for x, y in [(2, 0), (2, 1), (0, 2), (1, 2)]:
        b.place(WHITE, x, y)
    assert [b.at(0, 0), b.at(0, 1), b.at(1, 0), b.at(1, 1)] == [0, 0, 0, 0]


def test_suicide_corner():
    b = Board()
    b.place(WHITE, 0, 0)
    b.place(BLACK, 1, 0)
    b.place(BLACK, 0, 1)
    assert b.at(0, 0) == 0
    assert not b.is_valid(WHITE, 0, 0)
    assert b.is_valid(BLACK, 0, 0)


def test_suicide_middle():
    b = Board()
    b.place(BLACK, 9, 9)
    for x, y in [(8, 9), (10, 9), (9, 8), (9, 10)]:
        b.place(WHITE, x, y)
    assert b.at(9, 9) == 0
    assert not b.is_valid(BLACK, 9, 9)
    assert b.is_valid(WHITE, 9, 9)


def test_ko():
    b = Board()
    b.place(BLACK, 0, 0)
    b.place(BLACK, 0, 2)
    b.place(BLACK, 1, 1)
    b.place(WHITE, 1, 0)
    b.place(WHITE, 0, 1)
    assert not b.is_valid(BLACK, 0, 0)


def test_double_liberty_subtraction():
    b = Board()
    for x, y in [(1, 1), (1, 2), (2, 1), (0, 2), (2, 0)]:
        b.place(BLACK, x, y)
    for x, y in [(0, 3), (3, 0), (1, 3), (3, 1), (2, 2)]:
        b.place(WHITE, x, y)
    assert b.is_valid(WHITE, 0, 1) and b.is_valid(WHITE, 1, 0)
    b.place(WHITE, 0, 1)
    assert b.at(0, 1) == WHITE
    assert all(b.at(x, y) == BLACK for x, y in [(1, 1), (1, 2), (2, 1), (0, 2), (2, 0)])


def test_turns():
    b = Board(0.5)
    assert b.to_move() == BLACK
    b.place(BLACK, 0, 0)
    assert b.to_move() == WHITE
    b.place(WHITE, 1, 1)
    assert b.to_move() == BLACK
    b.place(WHITE, 2, 2)
    assert b.to_move() == BLACK


# ---- libdg_go/board_fast.rs:548-562 ------------------------------------------------------------------------

def test_get_n_liberty_if():
    b = Board()
    b.place(BLACK, 0, 1)
    b.place(BLACK, 1, 0)
    b.place(WHITE, 1, 1)
    b.place(WHITE, 2, 0)
    assert b.get_n_liberty_if(WHITE, 0, 0) == 1
    assert b.get_n_liberty_if(WHITE, 2, 1) == 4
    assert b.get_n_liberty_if(WHITE, 0, 2) == 2
    assert b.get_n_liberty_if(WHITE, 9, 9) == 4
    b.place(WHITE, 0, 2)
    assert b.get_n_liberty_if(WHITE, 0, 0) == 2


def test_bench_position_liberty_if():   # board_fast.rs:564-581
    b = Board()
    for x, y in [(1, 0), (1, 1), (1, 2), (0, 2)]:
        b.place(BLACK, x, y)
    for x, y in [(0, 1), (2, 0), (2, 1), (2, 2), (0, 3), (1, 3)]:
        b.place(WHITE, x, y)
    assert b.get_n_liberty_if(WHITE, 0, 0) == 3


# ---- libdg_go/utils/ladder.rs:187-351 ----------------------------------------------------------------------

def test_ladder_corner_capture():
    b = Board()
    for x, y in [(0, 0), (0, 18), (18, 0), (18, 18)]:
        b.place(BLACK, x, y)
    want = {(1, 0), (0, 1), (18, 17), (17, 18), (1, 18), (18, 1), (0, 17), (17, 0)}
    for y in range(19):
        for x in range(19):
            if b.is_valid(WHITE, x, y):
                assert b.is_ladder_capture(WHITE, x, y) == ((x, y) in want), (x, y)


def test_ladder_capture():
    b = Board()
    b.place(WHITE, 3, 3)
    for x, y in [(2, 3), (3, 2), (4, 2)]:
        b.place(BLACK, x, y)
    for y in range(19):
        for x in range(19):
            if b.is_valid(BLACK, x, y):
                assert b.is_ladder_capture(BLACK, x, y) == ((x, y) == (3, 4)), (x, y)


def test_ladder_escape():
    b = Board()
    b.place(WHITE, 3, 3)
    b.place(WHITE, 15, 15)
    for x, y in [(2, 3), (3, 2), (4, 2), (3, 4)]:
        b.place(BLACK, x, y)
    for y in range(19):
        for x in range(19):
            if b.is_valid(WHITE, x, y):
                assert not b.is_ladder_capture(BLACK, x, y)
                assert b.is_ladder_escape(WHITE, x, y) == ((x, y) == (4, 3)), (x, y)


NOT_LADDER = [
    (1, 15, 3), (2, 3, 15), (1, 16, 15), (2, 3, 2), (1, 14, 16), (2, 2, 4), (1, 3, 5), (2, 3, 4),
    (1, 4, 5), (2, 4, 4), (1, 5, 5), (2, 1, 6), (1, 12, 2), (2, 16, 9), (1, 16, 7), (2, 16, 12),
    (1, 6, 3), (2, 14, 9), (1, 15, 10), (2, 15, 9), (1, 14, 7), (2, 16, 2), (1, 16, 3), (2, 15, 2),
    (1, 14, 2), (2, 14, 1), (1, 13, 1), (2, 14, 3), (1, 13, 2), (2, 17, 3), (1, 17, 4), (2, 17, 1),
    (1, 18, 3), (2, 17, 2), (1, 14, 4), (2, 15, 1), (1, 17, 11), (2, 17, 12), (1, 17, 10), (2, 15, 11),
    (1, 17, 9), (2, 14, 10), (1, 12, 8), (2, 5, 16), (1, 12, 10), (2, 14, 14), (1, 12, 15), (2, 12, 14),
    (1, 11, 14), (2, 12, 13), (1, 11, 15), (2, 15, 16), (1, 15, 15), (2, 13, 16), (1, 13, 15), (2, 14, 15),
    (1, 14, 17), (2, 15, 17), (1, 13, 17), (2, 17, 16), (1, 13, 14), (2, 15, 14), (1, 13, 13), (2, 14, 12),
    (1, 5, 2), (2, 11, 9), (1, 11, 8), (2, 8, 8), (1, 8, 6), (2, 4, 1), (1, 1, 14), (2, 1, 15),
    (1, 2, 15), (2, 2, 14), (1, 2, 16), (2, 1, 16), (1, 3, 14), (2, 2, 13), (1, 3, 16), (2, 4, 15),
    (1, 1, 13), (2, 1, 17), (1, 2, 12), (2, 3, 13), (1, 4, 12), (2, 3, 12), (1, 3, 11),
]


def test_not_ladder_real_game():
    b = Board()
    for c, x, y in NOT_LADDER:
        b.place(c, x, y)
    assert b.is_ladder_escape(WHITE, 4, 13)


@pytest.mark.parametrize("first", [(1, 2), (3, 4)])
def test_not_ladder_due_to_self_atari(first):
    b = Board()
    for c, (x, y) in [(BLACK, first), (WHITE, (2, 4)), (BLACK, (2, 3)), (WHITE, (1, 5)), (BLACK, (1, 4))]:
        b.place(c, x, y)
    assert not b.is_ladder_capture(WHITE, 1, 3)


# ---- libdg_go/utils/symmetry.rs:148-201, libdg_mcts/predictor.rs:99-107 ---------------------------------------

@pytest.mark.parametrize("t", range(8))
def test_symmetry_bijection(t):
    seen = {go.symmetry_apply(t, i) for i in range(361)}
    assert seen == set(range(361))
    inv = go.lib().dgo_symmetry_inverse(t)
    assert all(go.symmetry_apply(inv, go.symmetry_apply(t, i)) == i for i in range(361))


def test_rot180_maps_0_to_360():
    assert go.symmetry_apply(go.ROT180, 0) == 360
    assert go.symmetry_apply(go.ROT180, 361) == 361


def test_point_packed_index():   # point.rs:228-243
    assert (72 % 19, 72 // 19) == (15, 3)


# ---- dg_tests/tests/real_games.rs:49,74,117 ------------------------------------------------------------------

@pytest.mark.parametrize("i", range(3))
def test_real_game_hashes(i):
    z = np.load(go._GOLDEN)
    out = go.replay(z[f"kat{i}_colors"], z[f"kat{i}_moves"], hashes=True)
    assert int(out["hash"][-1]) == int(z[f"kat{i}_hash"][0])


def test_example_games_replay_legally():   # dg_tests/fixtures/example_games.sgf through common/mod.rs:54-64
    games = go.load_games()
    assert len(games) == 99
    total = 0
    for colors, moves, komi in games:
        go.replay(colors, moves, komi, hashes=True)      # raises on an illegal move
        total += len(moves)
    assert total == 18649


# ---- features (features.rs:154-250); values are not pinned by the reference, pinned here from its doc-comment ----

def test_features_shape_and_empty_board():   # features.rs:477-493 asserts only the length
    b = Board(7.5)
    f = b.features(BLACK).astype(np.float32)
    assert f.shape == (361, 32)
    assert (f[:, 0] == 1.0).all() and (f[:, 1] == 0).all() and (f[:, 2] == 0).all()
    # empty board: every point is a legal move for both with min(4, neighbours) liberties
    assert (f[:, 11] == 1).all() and (f[:, 23] == 1).all()
    assert f[go.idx(0, 0), 12] == 1 and f[go.idx(0, 0), 13] == 0          # corner: 2 liberties
    assert f[go.idx(9, 9), 14] == 1 and f[go.idx(9, 9), 15] == 0          # centre: 4
    assert f[:, 3:11].sum() == 0 and f[:, 17:23].sum() == 0 and f[:, 29:].sum() == 0


def test_features_komi_plane():
    for komi, want in [(7.5, 1.0), (0.0, 0.5), (-7.5, 0.0), (0.5, np.float16(0.5 + 0.25 / 7.5)), (20.0, 1.0)]:
        f = Board(komi).features(WHITE)
        assert (f[:, 1] == np.float16(want)).all() and (f[:, 0] == 0).all()


def test_features_history_liberties_ko_ladder():
    b = Board()
    b.place(BLACK, 0, 0)
    b.place(BLACK, 0, 2)
    b.place(BLACK, 1, 1)
    b.place(WHITE, 1, 0)
    b.place(WHITE, 0, 1)            # captures (0,0): ko for black at (0,0)
    f = b.features(BLACK).astype(np.float32)
    assert f[go.idx(0, 1), 3] == 1 and f[:, 3].sum() == 1          # most recent move
    assert f[go.idx(1, 0), 4] == 1 and f[:, 4].sum() == 1          # the one before
    assert (f[:, 2] == 1).all()                                    # some move is super-ko
    assert f[go.idx(0, 0), 29] == 1 and f[:, 29].sum() == 1
    assert f[go.idx(1, 1), 5:11].tolist() == [1, 1, 0, 0, 0, 0]    # own stone with 2 liberties
    assert f[go.idx(0, 1), 17:23].tolist() == [1, 0, 0, 0, 0, 0]   # opponent stone in atari
    assert f[go.idx(0, 0), 11:17].tolist() == [1, 0, 0, 0, 0, 0]   # black retaking: 1 liberty after capture
    # white connecting at (0,0) leaves the three stones with the single liberty (2,0)
    assert f[go.idx(0, 0), 23:29].tolist() == [1, 0, 0, 0, 0, 0]


def test_features_ladder_planes():
    b = Board()
    b.place(WHITE, 3, 3)
    for x, y in [(2, 3), (3, 2), (4, 2)]:
        b.place(BLACK, x, y)
    f = b.features(BLACK).astype(np.float32)
    assert f[go.idx(3, 4), 30] == 1 and f[:, 30].sum() == 1
    b.place(WHITE, 15, 15)
    b.place(BLACK, 3, 4)
    f = b.features(WHITE).astype(np.float32)
    assert f[go.idx(4, 3), 31] == 1 and f[:, 31].sum() == 1


@pytest.mark.parametrize("t", range(8))
def test_features_symmetry_is_a_permutation(t):
    colors, moves, komi = go.load_games()[3]
    b = Board(komi)
    for c, m in zip(colors[:120], moves[:120]):
        if m < 361:
            b.place_index(int(c), int(m))
    base = b.features(BLACK, go.IDENTITY)
    sym = b.features(BLACK, t)
    perm = np.array([go.symmetry_apply(t, i) for i in range(361)])
    assert (sym[perm] == base).all()


# ---- utils/benson.rs:616-677, utils/score.rs:289-349, libdg_mcts/options.rs:217-262 -------------------------------

def test_benson_empty_is_all_valid():
    b = Board(0.5)
    assert (b.benson(BLACK) == 0).all() and (b.benson(WHITE) == 0).all()


def test_benson_eyes():
    b = Board(0.5)
    for x, y in [(0, 1), (1, 1), (2, 0), (2, 1), (3, 1), (4, 0), (4, 1)]:
        b.place(WHITE, x, y)
    b.place(BLACK, 0, 0)
    st = b.benson(WHITE)
    for y in range(19):
        for x in range(19):
            if b.at(x, y) == WHITE:
                assert st[go.idx(x, y)] == 1, (x, y)
    for x, y in [(0, 0), (1, 0), (3, 0)]:
        assert st[go.idx(x, y)] == 2, (x, y)


BENSON_BENCH = [
    (1, 15, 16), (2, 16, 3), (1, 3, 2), (2, 3, 15), (1, 14, 3), (2, 14, 2), (1, 13, 2), (2, 15, 2),
    (1, 12, 3), (2, 16, 5), (1, 9, 2), (2, 2, 5), (1, 3, 4), (2, 1, 3), (1, 3, 5), (2, 2, 6),
    (1, 3, 6), (2, 2, 7), (1, 4, 8), (2, 2, 2), (1, 2, 1), (2, 3, 9), (1, 4, 9), (2, 3, 10),
    (1, 4, 10), (2, 3, 11), (1, 2, 16), (2, 2, 15), (1, 3, 16), (2, 5, 16), (1, 5, 17), (2, 9, 6),
    (1, 16, 13), (2, 8, 3), (1, 8, 2), (2, 7, 3), (1, 6, 16), (2, 5, 15), (1, 7, 17), (2, 9, 3),
    (1, 10, 2), (2, 13, 16), (1, 11, 16), (2, 15, 14), (1, 16, 14), (2, 15, 17), (1, 16, 16), (2, 13, 14),
    (1, 11, 14), (2, 16, 17), (1, 17, 17), (2, 14, 16), (1, 16, 18), (2, 14, 18), (1, 14, 12), (2, 15, 15),
    (1, 16, 15), (2, 16, 8), (1, 12, 17), (2, 13, 17), (1, 17, 18), (2, 12, 16), (1, 11, 17), (2, 11, 13),
    (1, 10, 13), (2, 11, 12), (1, 10, 12), (2, 12, 14), (1, 10, 14), (2, 11, 11), (1, 11, 9), (2, 11, 6),
    (1, 9, 9), (2, 13, 4), (1, 12, 4), (2, 13, 5), (1, 10, 5), (2, 10, 6), (1, 9, 5), (2, 8, 5),
    (1, 8, 6), (2, 7, 5), (1, 8, 7), (2, 7, 2), (1, 11, 5), (2, 12, 5), (1, 10, 3), (2, 4, 1),
    (1, 4, 2), (2, 5, 2), (1, 5, 3), (2, 6, 1), (1, 3, 1), (2, 6, 3), (1, 5, 4), (2, 13, 1),
    (1, 12, 1), (2, 13, 3), (1, 12, 2), (2, 8, 1),
]


def test_benson_midgame_position_has_nothing_settled():   # benson.rs:585-613 (bench body asserts is_valid everywhere)
    b = Board(0.5)
    for c, x, y in BENSON_BENCH:
        b.place(c, x, y)
    assert (b.benson(WHITE) == 0).all()


def test_is_scorable_kats():
    b = Board(7.5)
    b.place(BLACK, 0, 0)
    assert not b.is_scorable()
    b = Board(7.5)
    for x, y in [(1, 0), (0, 1), (1, 1), (1, 2), (0, 3), (1, 3)]:
        b.place(WHITE, x, y)
    for x, y in [(2, 0), (2, 1), (2, 2), (2, 3), (0, 4), (1, 4), (2, 4)]:
        b.place(BLACK, x, y)
    assert not b.is_scorable()
    for color in (BLACK, WHITE):
        b = Board(0.5)
        for y in range(19):
            for x in range(1, 19, 2):
                b.place(color, x, y)
        assert b.is_scorable()


def test_simple_eye_kats():   # libdg_mcts/options.rs:217-262
    b = Board(0.5)
    for x, y in [(1, 0), (0, 1), (1, 1)]:
        b.place(BLACK, x, y)
    assert b.is_simple_eye(BLACK, 0, 0) and not b.is_simple_eye(WHITE, 0, 0)
    b = Board(0.5)
    for x, y in [(0, 0), (0, 1), (1, 1), (2, 1), (2, 0)]:
        b.place(BLACK, x, y)
    assert b.is_simple_eye(BLACK, 1, 0) and not b.is_simple_eye(WHITE, 1, 0)
    b = Board(0.5)
    for x, y in [(0, 1), (0, 2), (1, 0), (2, 0), (2, 2), (2, 1), (1, 2)]:
        b.place(BLACK, x, y)
    assert b.is_simple_eye(BLACK, 1, 1) and not b.is_simple_eye(WHITE, 1, 1)
    b.place(BLACK, 0, 0)
    assert b.is_simple_eye(BLACK, 1, 1) and not b.is_simple_eye(WHITE, 1, 1)


def test_eyes_should_be_territory():          # utils/score.rs:351-405
    b = Board(0.5)
    for x, y in [(0, 1), (1, 1), (2, 0), (2, 1), (3, 1), (4, 0), (4, 1)]:
        b.place(WHITE, x, y)
    b.place(BLACK, 0, 0)
    b.place(BLACK, 9, 9)
    t = b.territory()
    for y in range(19):
        for x in range(19):
            want = 1 if (x, y) == (9, 9) else 2      # alive white stones, their eyes, the dead stone at (0,0), everything they reach;
            assert t[go.idx(x, y)] == want, (x, y)   # the lone black stone counts for black ("seki")
    assert b.result() == "W+359.5"             # 360 points + 0.5 komi against the one black stone


def test_result_of_score_kats():              # utils/score.rs:289-330 positions through game_result.rs:78-93
    b = Board(7.5)
    b.place(BLACK, 0, 0)
    assert b.result() == "W+6.5"              # one stone vs komi: nothing is unconditionally alive
