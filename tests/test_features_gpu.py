"""Device feature kernel (csrc/features.cu): V1 planes + legal moves from raw stones, bit for bit against the host
code and the oracle restatement of the reference (features.rs:154-250, board.rs:151-153) -- every ply of the fixture
games, capture-heavy random playouts, all 8 symmetries; and the whole network from raw positions against the
compact-position path."""
import numpy as np
import pytest

from dream_go_b200 import go as pgo, nn, weights
from oracle import go as ogo

pytestmark = pytest.mark.gpu
BLACK, WHITE = 1, 2


@pytest.fixture(scope="module")
def engine(small_net):
    net = nn.Network.from_tensors(small_net, max_batch=512, num_workspaces=2)
    yield net
    net.close()


DEVICE_LADDERS = 0x08      # DG_RAW_DEVICE_LADDERS


def replay_raw(colors, moves, komi, symmetries=None, device_ladders=False):
    """Raw positions, host planes and host legal masks for every ply of a game (to_move = the move's colour).
    device_ladders: the raw positions carry no ladder masks, the device has to read the ladders itself."""
    board = pgo.Board(komi)
    raws, planes, legal = [], [], []
    for i, (c, m) in enumerate(zip(colors, moves)):
        s = 0 if symmetries is None else int(symmetries[i])
        raws.append(board.raw_position(int(c), s | (DEVICE_LADDERS if device_ladders else 0))[0])
        if device_ladders:
            assert not raws[-1]["ladder_capture"].any() and not raws[-1]["ladder_escape"].any()
        p, lg = board.features_packed(int(c), s, legal=True)
        planes.append(p[0])
        legal.append(lg)
        if m < 361:
            board.place_index(int(c), int(m))
    return np.array(raws, nn.RAW_DTYPE), np.array(planes, nn.PACKED_DTYPE), np.array(legal)


def check(engine, raws, planes, legal):
    for at in range(0, len(raws), 512):
        got_planes, got_legal = engine.features_raw(raws[at:at + 512])
        want = planes[at:at + 512]
        bad = np.argwhere(got_planes["planes"] != want["planes"])
        assert len(bad) == 0, f"first mismatch (position, point) = {bad[0]}: {got_planes['planes'][tuple(bad[0])]:#x} != {want['planes'][tuple(bad[0])]:#x}"
        assert (got_planes["k_bits"] == want["k_bits"]).all()
        assert (got_legal == legal[at:at + 512]).all()


def test_fixture_games_bit_exact(engine):
    ogo.use_default_zobrist()
    games = ogo.load_games()
    n = 0
    for colors, moves, komi in games[::4]:
        raws, planes, legal = replay_raw(colors, moves, komi)
        check(engine, raws, planes, legal)
        n += len(raws)
    # and the host planes are the oracle's (tests/test_go_parity.py does all 99 games on the CPU)
    colors, moves, komi = games[0]
    raws, planes, legal = replay_raw(colors, moves, komi)
    want = ogo.replay(colors, moves, komi, features=True, legal=True)
    got_planes, got_legal = engine.features_raw(raws[:512])
    assert (pgo.unpack_features(got_planes).view(np.uint16) == want["features"][:512].view(np.uint16)).all()
    assert (got_legal == want["legal"][:512]).all()
    assert n > 4000


@pytest.mark.parametrize("seed", range(4))
def test_random_playouts_and_symmetries_bit_exact(engine, seed):
    from test_go_parity import random_playout
    colors, moves = random_playout(977 + seed, 420, komi=6.5)
    rng = np.random.default_rng(seed)
    raws, planes, legal = replay_raw(colors, moves, 6.5, symmetries=rng.integers(0, 8, size=len(moves)))
    check(engine, raws, planes, legal)


def test_ko_and_capture_positions(engine):
    b = pgo.Board(7.5)
    for c, x, y in [(1, 0, 0), (1, 0, 2), (1, 1, 1), (2, 1, 0), (2, 0, 1)]:     # board.rs:341-351: ko at (0, 0)
        b.place(c, x, y)
    raws = np.concatenate([b.raw_position(BLACK, t) for t in range(8)] + [b.raw_position(WHITE, t) for t in range(8)])
    planes = np.concatenate([b.features_packed(BLACK, t) for t in range(8)] + [b.features_packed(WHITE, t) for t in range(8)])
    legal = np.stack([b.legal_moves(BLACK)] * 8 + [b.legal_moves(WHITE)] * 8)
    check(engine, raws, planes, legal)
    got, _ = engine.features_raw(raws[:1])
    assert (got["planes"][0] & 4).all() and got["planes"][0][0] & (1 << 29)     # the ko planes are really set


def test_network_from_raw_positions_matches_compact_path(engine):
    colors, moves, komi = ogo.load_games()[5]
    raws, planes, legal = replay_raw(colors[:96], moves[:96], komi)
    out, got_legal = engine.forward_raw(raws)
    want = engine.forward_packed(planes)
    assert (out.value.view(np.uint16) == want.value.view(np.uint16)).all()
    assert (out.policy.view(np.uint16) == want.policy.view(np.uint16)).all()
    assert (got_legal == legal).all()
    # a smaller batch right after a larger one: no stale rows may survive behind the last position
    out, _ = engine.forward_raw(raws[:5])
    want = engine.forward_packed(planes[:5])
    assert (out.value.view(np.uint16) == want.value.view(np.uint16)).all()
    assert (out.policy.view(np.uint16) == want.policy.view(np.uint16)).all()


def test_long_random_game_bit_exact(engine):
    """1,500 plies of random play (board fills, big captures, re-play on freed points, hash ring wrapping)."""
    from test_go_parity import random_playout
    colors, moves = random_playout(5, 1500, pass_rate=0.0)
    raws, planes, legal = replay_raw(colors, moves, 7.5)
    check(engine, raws, planes, legal)


# ---- ladder reading on the device (utils/ladder.rs:53-179; csrc/features.cu: namespace lad) -----------------------------------

def test_device_ladder_reader_on_the_reference_ladder_positions(engine):
    """The positions of the reference's own ladder tests (utils/ladder.rs:187-351): the device reads planes 30 / 31 itself
    (one warp per reading) and they are the host's, which tests/test_go_parity.py pins to the oracle."""
    from test_oracle_go import NOT_LADDER
    boards = []
    b = pgo.Board(7.5)                                   # ladder.rs: corner captures
    for x, y in [(0, 0), (0, 18), (18, 0), (18, 18)]:
        b.place(BLACK, x, y)
    boards.append((b, WHITE, {"capture": 8}))
    b = pgo.Board(7.5)                                   # a working ladder for black at (3, 4)
    b.place(WHITE, 3, 3)
    for x, y in [(2, 3), (3, 2), (4, 2)]:
        b.place(BLACK, x, y)
    boards.append((b, BLACK, {"capture": 1}))
    b = pgo.Board(7.5)                                   # ... and the escape at (4, 3) once a breaker stands at (15, 15)
    b.place(WHITE, 3, 3)
    b.place(WHITE, 15, 15)
    for x, y in [(2, 3), (3, 2), (4, 2), (3, 4)]:
        b.place(BLACK, x, y)
    boards.append((b, WHITE, {"escape": 1}))
    b = pgo.Board(7.5)                                   # real-game position: (4, 13) escapes
    for c, x, y in NOT_LADDER:
        b.place(c, x, y)
    boards.append((b, WHITE, {"escape_at": (4, 13)}))
    for first in [(1, 2), (3, 4)]:                       # not a ladder because of self-atari
        b = pgo.Board(7.5)
        for c, (x, y) in [(BLACK, first), (WHITE, (2, 4)), (BLACK, (2, 3)), (WHITE, (1, 5)), (BLACK, (1, 4))]:
            b.place(c, x, y)
        boards.append((b, WHITE, {"no_capture_at": (1, 3)}))
    for b, tm, expect in boards:
        for sym in (0, 5):
            raw = b.raw_position(tm, sym | DEVICE_LADDERS)
            want, want_legal = b.features_packed(tm, sym, legal=True)
            got, got_legal = engine.features_raw(raw)
            assert (got["planes"] == want["planes"]).all() and (got_legal[0] == want_legal).all()
        ident = b.features_packed(tm, 0)["planes"][0]
        if "capture" in expect:
            assert int(((ident >> 30) & 1).sum()) == expect["capture"]
        if "escape" in expect:
            assert int(((ident >> 31) & 1).sum()) == expect["escape"]
        if "escape_at" in expect:
            x, y = expect["escape_at"]
            assert (ident[19 * y + x] >> 31) & 1
        if "no_capture_at" in expect:
            x, y = expect["no_capture_at"]
            assert not (ident[19 * y + x] >> 30) & 1


def test_device_ladder_reader_on_games_and_playouts(engine):
    """Every ply of fixture games and of capture-heavy random playouts, random symmetries: the device's own ladder planes
    are the host reader's, and so is everything else in the planes."""
    from test_go_parity import random_playout
    games = ogo.load_games()
    n_ladders = 0
    for k, (colors, moves, komi) in enumerate(games[1::6]):
        rng = np.random.default_rng(k)
        raws, planes, legal = replay_raw(colors, moves, komi, symmetries=rng.integers(0, 8, size=len(moves)), device_ladders=True)
        check(engine, raws, planes, legal)
        n_ladders += int(((planes["planes"] >> 30) != 0).sum())
    assert n_ladders > 200                               # the games do contain ladders
    for seed in range(3):
        colors, moves = random_playout(4242 + seed, 500, komi=7.5)
        raws, planes, legal = replay_raw(colors, moves, 7.5, device_ladders=True)
        check(engine, raws, planes, legal)


def test_long_ladder_across_the_board(engine):
    """A ladder that runs from one corner region across the whole board (30+ steps), with and without a breaker."""
    for breaker in (None, (15, 15), (14, 15), (16, 14)):
        b = pgo.Board(7.5)
        b.place(WHITE, 2, 2)
        for x, y in [(1, 2), (2, 1), (3, 1)]:
            b.place(BLACK, x, y)
        if breaker:
            b.place(WHITE, *breaker)
        for tm in (BLACK, WHITE):
            raw = b.raw_position(tm, DEVICE_LADDERS)
            want = b.features_packed(tm, 0)
            got, _ = engine.features_raw(raw)
            assert (got["planes"] == want["planes"]).all(), (breaker, tm)
        if breaker is None:
            assert (b.features_packed(BLACK, 0)["planes"][0][19 * 3 + 2] >> 30) & 1     # black (2, 3) starts the ladder


# ---- prior construction on the device (pool/policy_helper.rs) ------------------------------------------------------------

def check_priors(engine, boards, to_moves, searches, symmetries):
    raws = np.concatenate([b.raw_position(tm, s, search=k) for b, tm, k, s in zip(boards, to_moves, searches, symmetries)])
    out, legal, prior = engine.forward_raw_prior(raws)
    policy = out.policy.reshape(-1, 362)
    for i, (b, tm, k, s) in enumerate(zip(boards, to_moves, searches, symmetries)):
        want = b.prior(tm, policy[i], s, 1.0, search=k)
        assert (np.isfinite(prior[i]) == np.isfinite(want)).all(), (i, k, s, np.flatnonzero(np.isfinite(prior[i]) != np.isfinite(want)))
        f = np.isfinite(want)
        assert (prior[i][f].view(np.uint32) == want[f].view(np.uint32)).all(), (i, k, s)
        assert (legal[i] == b.legal_moves(tm)).all()
    return prior


def test_device_priors_bit_exact_on_games(engine):
    games = ogo.load_games()
    boards, tms, kinds, syms = [], [], [], []
    rng = np.random.default_rng(3)
    for g in range(0, 99, 7):
        colors, moves, komi = games[g]
        b = pgo.Board(komi)
        for ply, (c, m) in enumerate(zip(colors, moves)):
            if ply % 23 == 0 or ply == len(moves) - 1:
                boards.append(b.clone()); tms.append(int(c)); kinds.append(int(rng.integers(0, 2))); syms.append(int(rng.integers(0, 8)))
            if m < 361:
                b.place_index(int(c), int(m))
    check_priors(engine, boards[:512], tms[:512], kinds[:512], syms[:512])
    assert len(boards) > 100


def test_device_priors_symmetric_and_settled_positions(engine):
    from test_go_parity import settled_position
    boards, tms, kinds, syms = [], [], [], []
    empty = pgo.Board(7.5)
    tengen = pgo.Board(7.5); tengen.place(1, 9, 9)
    for b, tm in ((empty, 1), (tengen, 2)):                # full symmetry group: orbit folding (policy_helper.rs:54-72)
        for s in range(8):
            for k in (0, 1):
                boards.append(b); tms.append(tm); kinds.append(k); syms.append(s)
    for seed in range(8):                                  # alive groups, eyes, dead stones: Benson on the device
        b = pgo.Board(6.5)
        for c, x, y in settled_position(seed):
            if b.at(x, y) == 0 and b.is_valid(c, x, y):
                b.place(c, x, y)
        assert (b.benson(1) == 2).any() and (b.benson(2) == 2).any()
        for tm in (1, 2):
            for k in (0, 1):
                boards.append(b); tms.append(tm); kinds.append(k); syms.append((seed + tm) % 8)
    prior = check_priors(engine, boards, tms, kinds, syms)
    assert np.isfinite(prior[0][:361]).sum() == 55         # empty board, standard search: one candidate per orbit
