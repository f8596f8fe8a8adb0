"""Parity of the product's host Go code (csrc/go_board.h through include/dg_go.h) with the oracle restatement of
the reference's libdg_go: bit-exact stones, hashes, legal moves, liberties, ladders, feature planes and (to fp32
rounding) priors.  CPU only -- this half of the hot path runs on the host in the reference too."""
import numpy as np
import pytest

from dream_go_b200 import go as pgo
from oracle import go as ogo

BLACK, WHITE = 1, 2


@pytest.fixture(scope="module", autouse=True)
def _zobrist():
    ogo.use_default_zobrist()


def random_playout(seed: int, plies: int, komi: float = 7.5, pass_rate: float = 0.02):
    """Uniform-random legal playout (legality decided by the ORACLE), returns (colors, moves)."""
    rng = np.random.default_rng(seed)
    board = ogo.Board(komi)
    colors, moves = [], []
    color = BLACK
    for _ in range(plies):
        legal = np.flatnonzero(board.legal_mask(color))
        if len(legal) == 0 or rng.random() < pass_rate:
            colors.append(color)
            moves.append(361)
        else:
            m = int(rng.choice(legal))
            board.place_index(color, m)
            colors.append(color)
            moves.append(m)
        color = 3 - color
    return np.array(colors, np.uint8), np.array(moves, np.uint16)


def assert_same_replay(colors, moves, komi):
    want = ogo.replay(colors, moves, komi, features=True, legal=True, hashes=True)
    got = pgo.replay(colors, moves, komi, features=True, legal=True, hashes=True)
    assert (got["hash"] == want["hash"]).all()
    assert (got["legal"] == want["legal"]).all()
    feats = pgo.unpack_features(got["features"])
    bad = np.argwhere(feats.view(np.uint16) != want["features"].view(np.uint16))
    assert len(bad) == 0, f"first mismatch (ply, point, plane) = {bad[0]}"


# ---- the reference's unit KATs, run against the PRODUCT (same assertions as tests/test_oracle_go.py) -----------

def test_reference_unit_kats_on_product():
    b = pgo.Board()
    b.place(BLACK, 9, 9)
    for x, y in [(8, 9), (10, 9), (9, 8), (9, 10)]:
        b.place(WHITE, x, y)
    assert b.at(9, 9) == 0 and not b.is_valid(BLACK, 9, 9) and b.is_valid(WHITE, 9, 9)      # board.rs:281-337
    b = pgo.Board()
    for c, x, y in [(1, 0, 0), (1, 0, 2), (1, 1, 1), (2, 1, 0), (2, 0, 1)]:
        b.place(c, x, y)
    assert not b.is_valid(BLACK, 0, 0)                                                       # board.rs:341-351 (ko)
    b = pgo.Board()
    b.place(BLACK, 0, 1); b.place(BLACK, 1, 0); b.place(WHITE, 1, 1); b.place(WHITE, 2, 0)   # board_fast.rs:548-562
    assert [b.get_n_liberty_if(WHITE, *p) for p in [(0, 0), (2, 1), (0, 2), (9, 9)]] == [1, 4, 2, 4]
    b.place(WHITE, 0, 2)
    assert b.get_n_liberty_if(WHITE, 0, 0) == 2


def test_ladder_kats_on_product():   # utils/ladder.rs:187-351
    b = pgo.Board()
    for x, y in [(0, 0), (0, 18), (18, 0), (18, 18)]:
        b.place(BLACK, x, y)
    want = {(1, 0), (0, 1), (18, 17), (17, 18), (1, 18), (18, 1), (0, 17), (17, 0)}
    for y in range(19):
        for x in range(19):
            if b.is_valid(WHITE, x, y):
                assert b.is_ladder_capture(WHITE, x, y) == ((x, y) in want)
    b = pgo.Board()
    b.place(WHITE, 3, 3); b.place(WHITE, 15, 15)
    for x, y in [(2, 3), (3, 2), (4, 2), (3, 4)]:
        b.place(BLACK, x, y)
    for y in range(19):
        for x in range(19):
            if b.is_valid(WHITE, x, y):
                assert not b.is_ladder_capture(BLACK, x, y)
                assert b.is_ladder_escape(WHITE, x, y) == ((x, y) == (4, 3))
    from test_oracle_go import NOT_LADDER
    b = pgo.Board()
    for c, x, y in NOT_LADDER:
        b.place(c, x, y)
    assert b.is_ladder_escape(WHITE, 4, 13)


def test_raw_positions_against_the_oracle():
    """`dg_raw_position` (what the device feature kernel starts from): stones, visited points, hashes, last moves and the two
    ladder planes -- the only planes the host still computes on that path -- against the ORACLE's board and feature planes,
    on random play (weak chains everywhere: many ladder readings, most decided at the first step, some deep)."""
    def bits(words):
        return np.unpackbits(np.asarray(words, "<u4").view(np.uint8), bitorder="little")[:361]

    colors, moves = random_playout(77, 260)
    po, oo = pgo.Board(7.5), ogo.Board(7.5)
    checked = ladders = 0
    for ply, (c, m) in enumerate(zip(colors, moves)):
        if m < 361:
            po.place_index(int(c), int(m))
            oo.place_index(int(c), int(m))
        if ply % 3:
            continue
        to_move = 3 - int(c)
        raw = po.raw_position(to_move)[0]
        want = oo.features(to_move).reshape(361, 32).astype(np.float32)
        stones = po.stones()
        assert (bits(raw["black"]) == (stones == BLACK)).all() and (bits(raw["white"]) == (stones == WHITE)).all()
        assert (bits(raw["ladder_capture"]) == (want[:, 30] != 0)).all(), ply
        assert (bits(raw["ladder_escape"]) == (want[:, 31] != 0)).all(), ply
        assert int(raw["hash"]) == po.zobrist_hash() and int(raw["to_move"]) == to_move
        ladders += int(bits(raw["ladder_capture"]).sum() + bits(raw["ladder_escape"]).sum())
        checked += 1
    assert checked > 80 and ladders > 20


@pytest.mark.parametrize("t", range(8))
def test_symmetry_tables(t):
    assert [pgo.symmetry_apply(t, i) for i in range(362)] == [ogo.symmetry_apply(t, i) for i in range(362)]
    assert pgo.lib().dg_symmetry_inverse(t) == ogo.lib().dgo_symmetry_inverse(t)


# ---- bit-exact replay parity on the reference's corpus and on capture-heavy random games -------------------------

def test_example_games_bit_exact():
    """All 18,649 positions of dg_tests/fixtures/example_games.sgf: hash, legal mask and all 32 planes."""
    for colors, moves, komi in ogo.load_games():
        assert_same_replay(colors, moves, komi)


@pytest.mark.parametrize("seed", range(6))
def test_random_playouts_bit_exact(seed):
    """Random play fills the board and produces the captures, kos and snapbacks that real games rarely have."""
    colors, moves = random_playout(20261017 + seed, 420, komi=[7.5, 6.5, 0.5, -3.5, 0.0, 20.0][seed])
    assert_same_replay(colors, moves, [7.5, 6.5, 0.5, -3.5, 0.0, 20.0][seed])


def test_features_every_symmetry_and_fp16_form():
    colors, moves, komi = ogo.load_games()[7]
    po, oo = pgo.Board(komi), ogo.Board(komi)
    for i, (c, m) in enumerate(zip(colors[:150], moves[:150])):
        if m < 361:
            po.place_index(int(c), int(m))
            oo.place_index(int(c), int(m))
    for t in range(8):
        for to_move in (BLACK, WHITE):
            want = oo.features(to_move, t)
            assert (po.features(to_move, t).view(np.uint16) == want.view(np.uint16)).all()
            packed, legal = po.features_packed(to_move, t, legal=True)
            assert (pgo.unpack_features(packed)[0].view(np.uint16) == want.view(np.uint16)).all()
            assert (legal == oo.legal_mask(to_move)).all()
            assert (po.legal_moves(to_move) == legal).all()
    assert (po.stones() == oo.stones()).all()
    for t in range(8):
        assert po.is_symmetric(t) == oo.is_symmetric(t)


def test_extract_batch_matches_single_calls():
    games = ogo.load_games()
    boards, tm = [], []
    for g in range(32):
        colors, moves, komi = games[g]
        b = pgo.Board(komi)
        n = 20 + 5 * g
        for c, m in zip(colors[:n], moves[:n]):
            if m < 361:
                b.place_index(int(c), int(m))
        boards.append(b)
        tm.append(b.to_move())
    sym = np.arange(32) % 8
    out, legal = pgo.extract_batch(boards, tm, sym, legal=True, threads=4)
    for i, b in enumerate(boards):
        one, lg = b.features_packed(tm[i], int(sym[i]), legal=True)
        assert out[i].tobytes() == one[0].tobytes()
        assert (legal[i] == lg).all()


# ---- prior construction (pool/policy_helper.rs) -------------------------------------------------------------------

def oracle_prior(board: "ogo.Board", to_move: int, policy_f16: np.ndarray, symmetry: int, sum_to: float) -> np.ndarray:
    """Restatement of create_initial_policy (policy_helper.rs:28-75, StandardSearch options.rs:53-57),
    add_valid_candidates (:87-104) and normalize_policy (:113-134; lane order of asm/sum_finite.rs:23-57)."""
    prior = np.full(368, -np.inf, np.float32)
    legal = board.legal_mask(to_move)
    prior[:361][legal != 0] = 0.0
    prior[361] = 0.0
    syms = [t for t in range(8) if board.is_symmetric(t)]
    indices = np.zeros(362, np.int64)
    indices[361] = 361
    for i in range(361):
        target = min(ogo.symmetry_apply(t, i) for t in syms)
        indices[i] = target
        if i != target:
            prior[i] = -np.inf
    src = policy_f16.astype(np.float32)
    prior[361] += src[361]
    inv = ogo.lib().dgo_symmetry_inverse(symmetry)
    for i in range(361):
        j = indices[ogo.symmetry_apply(inv, i)]
        prior[j] = np.float32(prior[j] + src[i])
    lanes = np.zeros(8, np.float32)
    for i in range(368):
        if np.isfinite(prior[i]):
            lanes[i & 7] = np.float32(lanes[i & 7] + prior[i])
    total = np.float32(np.float32(np.float32(lanes[0] + lanes[1]) + np.float32(lanes[2] + lanes[3])) +
                       np.float32(np.float32(lanes[4] + lanes[5]) + np.float32(lanes[6] + lanes[7])))
    finite = np.isfinite(prior)
    if total < 1e-6:
        prior[finite] = np.float32(sum_to) / np.float32(finite.sum())
    else:
        recip = np.float32(1.0) / np.float32(total / np.float32(sum_to))
        prior = (prior * recip).astype(np.float32)
    return prior


@pytest.mark.parametrize("case", ["empty", "opening", "midgame", "ko"])
def test_prior_matches_policy_helper(case):
    rng = np.random.default_rng(5)
    po, oo = pgo.Board(7.5), ogo.Board(7.5)
    plays = {"empty": [], "opening": [(1, 9, 9)], "ko": [(1, 0, 0), (1, 0, 2), (1, 1, 1), (2, 1, 0), (2, 0, 1)]}.get(case)
    if plays is None:
        colors, moves, _ = ogo.load_games()[11]
        plays = [(int(c), int(m) % 19, int(m) // 19) for c, m in zip(colors[:140], moves[:140]) if m < 361]
    for c, x, y in plays:
        po.place(c, x, y)
        oo.place(c, x, y)
    to_move = po.to_move()
    for symmetry in range(8):
        logits = rng.normal(size=362).astype(np.float32)
        policy = (np.exp(logits) / np.exp(logits).sum()).astype(np.float16)
        for sum_to in (1.0, 0.125):
            want = oracle_prior(oo, to_move, policy, symmetry, sum_to)
            got = po.prior(to_move, policy, symmetry, sum_to)
            assert (np.isfinite(got) == np.isfinite(want)).all()
            f = np.isfinite(want)
            assert (got[f].view(np.uint32) == want[f].view(np.uint32)).all()        # same operation order: bit-exact
            assert abs(float(got[f].sum()) - sum_to) < 1e-4
    if case == "empty":      # symmetric board: only one representative per orbit survives (policy_helper.rs:54-72)
        assert np.isfinite(po.prior(BLACK, np.full(362, 1 / 362, np.float16))[:361]).sum() == 55
    # all-zero policy -> uniform over the candidates (policy_helper.rs:119-124)
    z = po.prior(to_move, np.zeros(362, np.float16))
    f = np.isfinite(z)
    assert np.allclose(z[f], 1.0 / f.sum())


def test_prediction_with_transform_kat():
    """predictor.rs:99-107: un-rotating a Rot180 policy moves entry 0 to 360 and keeps the pass entry."""
    b = pgo.Board(7.5)
    b.place(BLACK, 3, 2)            # break every symmetry so no folding happens
    policy = (np.arange(362) / 1000.0).astype(np.float16)
    raw = b.prior(WHITE, policy, pgo.ROT180, sum_to=float(policy.astype(np.float32).sum() - policy[360 - pgo.idx(3, 2)]))
    assert abs(raw[360] - float(policy[0])) < 1e-3 and abs(raw[361] - float(policy[361])) < 1e-3


# ---- unconditional life / scoring candidates (utils/benson.rs, utils/score.rs, libdg_mcts/options.rs) ---------------

def settled_position(seed: int):
    """Walled-off territories with two-eyed groups plus random filler: exercises alive blocks, vital regions,
    dead stones inside eyes and regions that fail to be vital."""
    rng = np.random.default_rng(seed)
    stones = []
    for x in range(19):
        stones.append((BLACK, x, 1))
        stones.append((WHITE, x, 17))
        if x not in (3, 9):
            stones.append((BLACK, x, 0))
        if x not in (5, 6, 14):                       # a two-point region (not vital) and a one-point eye
            stones.append((WHITE, x, 18))
    stones += [(WHITE, 9, 0)] if seed % 2 else []     # a dead stone inside an eye
    for _ in range(int(rng.integers(0, 60))):
        stones.append((int(rng.integers(1, 3)), int(rng.integers(0, 19)), int(rng.integers(2, 16))))
    return stones


@pytest.mark.parametrize("seed", range(8))
def test_benson_scorable_candidates_parity(seed):
    po, oo = pgo.Board(6.5), ogo.Board(6.5)
    for c, x, y in settled_position(seed):
        if oo.at(x, y) == 0 and oo.is_valid(c, x, y):
            po.place(c, x, y)
            oo.place(c, x, y)
    assert (po.stones() == oo.stones()).all()
    for color in (BLACK, WHITE):
        assert (po.benson(color) == oo.benson(color)).all()
        for kind in (0, 1):
            assert (po.policy_candidates(color, kind) == oo.policy_candidates(color, kind)).all()
    assert po.is_scorable() == oo.is_scorable()
    assert (oo.benson(BLACK) != 0).any() and (oo.benson(WHITE) != 0).any()


def test_benson_parity_on_corpus_endgames():
    """Final and late positions of the fixture games (nothing is pinned by the reference here beyond its KATs)."""
    n_settled = 0
    for colors, moves, komi in ogo.load_games()[::3]:
        po, oo = pgo.Board(komi), ogo.Board(komi)
        for i, (c, m) in enumerate(zip(colors, moves)):
            if m < 361:
                po.place_index(int(c), int(m))
                oo.place_index(int(c), int(m))
            if i % 40 == 39 or i == len(moves) - 1:
                for color in (BLACK, WHITE):
                    want = oo.benson(color)
                    n_settled += int((want != 0).sum())
                    assert (po.benson(color) == want).all()
                    assert (po.policy_candidates(color, 1) == oo.policy_candidates(color, 1)).all()
                assert po.is_scorable() == oo.is_scorable()
    assert n_settled > 0


def test_scorable_kats_on_product():   # utils/score.rs:289-349
    for color in (BLACK, WHITE):
        b = pgo.Board(0.5)
        for y in range(19):
            for x in range(1, 19, 2):
                b.place(color, x, y)
        assert b.is_scorable()
    b = pgo.Board(7.5)
    b.place(BLACK, 0, 0)
    assert not b.is_scorable()


def test_territory_parity():                  # utils/score.rs:148-195
    for seed in range(4):
        po, oo = pgo.Board(6.5), ogo.Board(6.5)
        for c, x, y in settled_position(seed):
            if oo.at(x, y) == 0 and oo.is_valid(c, x, y):
                po.place(c, x, y)
                oo.place(c, x, y)
        want = oo.territory()
        assert (po.territory() == want).all()
        assert (want == 1).any() and (want == 2).any()
    for colors, moves, komi in ogo.load_games()[::9]:
        po, oo = pgo.Board(komi), ogo.Board(komi)
        for c, m in zip(colors, moves):
            if m < 361:
                po.place_index(int(c), int(m))
                oo.place_index(int(c), int(m))
        assert (po.territory() == oo.territory()).all()


@pytest.mark.parametrize("seed", [5, 6])
def test_long_random_games_bit_exact(seed):
    """1,500 plies of random play: the board fills up, large groups die and the freed areas are played in again and again
    (visited points, the 16-entry hash ring wrapping ~90 times, occasional kos) -- far beyond what real games reach."""
    colors, moves = random_playout(seed, 1500, pass_rate=0.0)
    stones = ogo.replay(colors, moves, 7.5, features=True)["features"].astype(np.float32)[:, :, 5].sum(axis=1)
    assert (np.diff(stones) < -5).any()                     # big captures happen
    assert_same_replay(colors, moves, 7.5)


# ---- the reference's zobrist constants at the product boundary (dg_go_set_zobrist) ----------------------------------

def test_real_game_hashes_through_the_product_board():
    """dg_tests/tests/real_games.rs:49,74,117: with `zobrist::TABLE` (zobrist.rs:18) loaded through dg_go_set_zobrist the
    PRODUCT board ends the three pinned games on the reference's hashes; afterwards the built-in table is back."""
    import numpy as np
    from dream_go_b200 import go as pgo
    from oracle import go as ogo
    z = np.load(ogo._GOLDEN)
    before = pgo.replay(z["kat0_colors"], z["kat0_moves"], hashes=True)["hash"][-1]
    pgo.set_zobrist(z["zobrist"])
    try:
        for i in range(3):
            out = pgo.replay(z[f"kat{i}_colors"], z[f"kat{i}_moves"], hashes=True)
            assert int(out["hash"][-1]) == int(z[f"kat{i}_hash"][0])
    finally:
        pgo.set_zobrist(None)
    assert pgo.replay(z["kat0_colors"], z["kat0_moves"], hashes=True)["hash"][-1] == before != int(z["kat0_hash"][0])


def test_out_of_range_arguments_are_rejected_at_the_boundary():
    from dream_go_b200 import go as pgo
    L = pgo.lib()
    b = pgo.Board(7.5)
    b.place(1, 3, 3)
    h = b.zobrist_hash()
    for point in (361, 362, -1, 100000):
        L.dg_board_place(b._h, 2, point)                       # pass / garbage: not played, nothing corrupted
        assert L.dg_board_is_valid(b._h, 2, point) == 0 and L.dg_board_at(b._h, point) == -1
        assert L.dg_board_get_n_liberty_if(b._h, 2, point) == -1
    L.dg_board_place(b._h, 3, 5)                               # not a colour
    assert b.zobrist_hash() == h and b.count() == 1 and b.to_move() == 2
    assert L.dg_symmetry_apply(8, 0) == -1 and L.dg_symmetry_apply(0, 362) == -1 and L.dg_symmetry_inverse(-1) == -1
    assert L.dg_symmetry_apply(3, 361) == 361
    # array-filling entry points: zeros (priors: -inf) instead of reading tables out of bounds
    import ctypes as C
    from dream_go_b200 import nn
    packed = np.ones(1, nn.PACKED_DTYPE)
    legal = np.ones(361, np.uint8)
    for to_move, symmetry in ((0, 0), (3, 0), (1, 8), (2, -1), (200, 200)):
        packed["planes"][:] = 7
        legal[:] = 1
        L.dg_board_features_packed(b._h, to_move, symmetry, packed.ctypes.data, legal.ctypes.data)
        assert not packed["planes"].any() and not legal.any()
        f16 = np.ones((361, 32), np.float16)
        L.dg_board_features_f16(b._h, to_move, symmetry, f16.ctypes.data)
        assert not f16.any()
        prior = np.zeros(368, np.float32)
        policy = np.full(362, 1 / 362, np.float16)
        L.dg_board_prior(b._h, to_move, 0, None, policy.ctypes.data, symmetry, C.c_float(1.0), prior.ctypes.data)
        assert np.isneginf(prior).all()
    raw = np.ones(1, nn.RAW_DTYPE)
    raw["black"][:] = 5
    L.dg_board_raw_position(b._h, 0, 0, raw.ctypes.data)
    assert not raw["black"].any() and int(raw["to_move"][0]) == 0
    out = np.ones(362, np.uint8)
    for to_move, search in ((0, 0), (1, 2), (2, -1)):
        out[:] = 1
        L.dg_board_policy_candidates(b._h, to_move, search, None, out.ctypes.data)
        assert not out.any()
    prior = np.zeros(368, np.float32)
    L.dg_board_prior(b._h, 1, 5, None, policy.ctypes.data, 0, C.c_float(1.0), prior.ctypes.data)
    assert np.isneginf(prior).all()
    alive = np.ones(361, np.uint8)
    L.dg_board_benson(b._h, 0, alive.ctypes.data)
    assert not alive.any()
    # a replay stops at a ply whose colour is not a colour, as at an illegal move
    colors = np.array([1, 2, 3, 1], np.uint8)
    moves = np.array([0, 1, 2, 3], np.uint16)
    assert L.dg_go_replay(C.c_float(7.5), colors.ctypes.data, moves.ctypes.data, 4, None, None, None) == -3
    # a batch with one bad entry: that entry comes back empty, the others are computed
    boards = (C.c_void_p * 2)(b._h, b._h)
    tm = np.array([2, 9], np.uint8)
    sym = np.array([0, 0], np.uint8)
    two = np.zeros(2, nn.PACKED_DTYPE)
    L.dg_go_extract_batch(boards, tm.ctypes.data, sym.ctypes.data, 2, two.ctypes.data, None, 1)
    assert two["planes"][0].any() and not two["planes"][1].any()
    assert b.zobrist_hash() == h and b.count() == 1


def test_board_copy_and_komi_through_the_boundary():
    """`dg_board_copy` / `dg_board_set_komi`: a copy is the same position (stones, hash, super-ko history, features) and is
    independent of its source afterwards; the komi only moves the two colour planes' constant."""
    colors, moves = random_playout(4, 120)
    src = pgo.Board(7.5)
    for c, m in zip(colors, moves):
        if m < 361:
            src.place_index(int(c), int(m))
    dst = pgo.Board(0.5)
    dst.copy_from(src)
    to_move = src.to_move()
    assert dst.zobrist_hash() == src.zobrist_hash() and dst.komi() == 7.5 and (dst.stones() == src.stones()).all()
    assert (dst.legal_moves(to_move) == src.legal_moves(to_move)).all()            # super-ko history included
    assert (dst.features(to_move).view(np.uint16) == src.features(to_move).view(np.uint16)).all()
    free = int(np.flatnonzero(src.legal_moves(to_move))[0])
    dst.place_index(to_move, free)
    assert dst.zobrist_hash() != src.zobrist_hash() and src.at(free % 19, free // 19) == 0
    src.set_komi(-3.5)
    oo = ogo.Board(-3.5)
    for c, m in zip(colors, moves):
        if m < 361:
            oo.place_index(int(c), int(m))
    assert src.komi() == -3.5
    assert (src.features(to_move).view(np.uint16) == oo.features(to_move).reshape(361, 32).view(np.uint16)).all()
