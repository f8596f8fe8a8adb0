"""Parity of the product's tree search (csrc/search*.h through include/dg_mcts.h) with the oracle restatement of
the reference's `dg_mcts`: under the same deterministic stub predictor, injected noise / leaf symmetries / random
numbers, the visit count of EVERY root child, the value estimates, the chosen move and the number of evaluated
positions are bit-identical -- with one probe in flight and with several.  CPU only (the predictor is a stub)."""
import numpy as np
import pytest

from dream_go_b200 import go as pgo, mcts as pm
from oracle import go as ogo, mcts as om
from mcts_common import dirichlet_sample, fake_predictor, hash_predictor, nan_predictor, uniform_predictor

BLACK, WHITE = 1, 2


def boards(plays, komi=7.5):
    po, oo = pgo.Board(komi), ogo.Board(komi)
    for c, m in plays:
        po.place_index(int(c), int(m))
        oo.place_index(int(c), int(m))
    return po, oo


def corpus_position(game: int, plies: int):
    colors, moves, komi = ogo.load_games()[game]
    return [(c, m) for c, m in zip(colors[:plies], moves[:plies]) if m < 361], komi


def assert_same_search(stub, po, oo, color, **kw):
    want_v, want_i, want_root, want_evals = om.predict(stub, oo, color, **kw)
    got_v, got_i, tree, got_evals = pm.predict(pm.python_predictor(stub), po, color, **kw)
    count, value, prior = tree.children()
    assert (count == want_root.count[:362]).all(), np.flatnonzero(count != want_root.count[:362])
    assert tree.total_count == want_root.total_count
    assert (prior.view(np.uint32) == want_root.prior[:362].view(np.uint32)).all()
    visited = count > 0
    assert (value[visited].view(np.uint32) == want_root.value[:362][visited].view(np.uint32)).all()
    assert got_i == want_i
    assert np.float32(got_v).view(np.uint32) == np.float32(want_v).view(np.uint32)
    assert got_evals == want_evals
    return tree, want_root


@pytest.mark.parametrize("probes", [1, 4])
def test_empty_board_deterministic(probes):
    po, oo = boards([])
    assert_same_search(hash_predictor(), po, oo, BLACK, deterministic=True, num_rollout=120, probes_per_round=probes,
                       leaf_symmetries=[0, 5, 3, 6, 1, 7, 2, 4])


@pytest.mark.parametrize("probes", [1, 3])
def test_midgame_with_noise_and_temperature(probes):
    plays, komi = corpus_position(5, 6)          # count < 8: the stochastic move choice is active
    po, oo = boards(plays, komi)
    color = po.to_move()
    _, policy, _ = om.full_forward(hash_predictor(), 0, oo, color)
    noise = dirichlet_sample(3, policy[:362])
    assert_same_search(hash_predictor(), po, oo, color, deterministic=False, num_rollout=100, probes_per_round=probes,
                       noise=noise, leaf_symmetries=[2, 2, 7, 0, 4], choose_at=0.37)


def tied_predictor():
    """Near-uniform policy with only three distinct values (ties in every block of 8) and values close to 0: what a
    random-init network gives -- the root ends up with hundreds of visited children."""
    import zlib

    def fn(feats: np.ndarray):
        n = feats.shape[0]
        v = np.empty(n, np.float16)
        p = np.empty((n, 362), np.float16)
        for i in range(n):
            rng = np.random.default_rng(zlib.crc32(np.ascontiguousarray(feats[i]).tobytes()))
            p[i] = (rng.integers(1, 4, size=362) / (2 * 362)).astype(np.float16)
            v[i] = np.float16(rng.normal() * 0.03)
        return v, p
    return fn


@pytest.mark.parametrize("game,plies,search,rollouts,min_children", [(5, 6, 0, 700, 257), (11, 150, 1, 300, 129)])
def test_wide_roots_with_tied_priors(game, plies, search, rollouts, min_children):
    """Hundreds of child records at the root (the parallel arrays grow 16 -> 512, eight records are scored at a time),
    equal priors everywhere (the argmax tie rule and the lazy candidate order decide every probe)."""
    plays, komi = corpus_position(game, plies)
    po, oo = boards(plays, komi)
    tree, _ = assert_same_search(tied_predictor(), po, oo, po.to_move(), search=search, deterministic=True, num_rollout=rollouts,
                                 probes_per_round=8, leaf_symmetries=[1, 6, 3, 0, 5])
    count, _, _ = tree.children()
    assert (count > 0).sum() >= min_children


@pytest.mark.parametrize("rollouts,probes", [(800, 8), (1700, 8), (900, 1)])
def test_full_rollout_budget_past_the_interpolation_knots(rollouts, probes):
    """Every BASELINE config searches with 800 rollouts: n = total + virtual visits then passes the knots at 800 and 1600 of
    the piecewise-linear UCT_EXP / FPU_REDUCE tables (src/libdg_utils/config.rs:181-195, 297-336) -- the second and
    third segment of the interpolation, which the short searches above never reach."""
    plays, komi = corpus_position(7, 30)
    po, oo = boards(plays, komi)
    tree, want_root = assert_same_search(hash_predictor(1.0), po, oo, po.to_move(), deterministic=True, num_rollout=rollouts,
                                         probes_per_round=probes, leaf_symmetries=[3, 0, 6, 1, 4, 7, 2, 5])
    assert tree.total_count + 32 * probes > 800          # n = visits + virtual visits really crossed the first knot
    if rollouts > 1600:
        assert tree.total_count > 1600                   # ... and the second


def test_late_game_scoring_search_and_tree_reuse():
    plays, komi = corpus_position(11, 150)
    po, oo = boards(plays, komi)
    color = po.to_move()
    kw = dict(search=1, deterministic=True, num_rollout=90, probes_per_round=2, leaf_symmetries=[1, 6, 3])
    tree, want_root = assert_same_search(hash_predictor(1.0), po, oo, color, **kw)
    # play the move, hand the sub-tree to the next search (self_play.rs:339-341, lib.rs:170-182)
    count, _, _ = tree.children()
    move = int(np.argmax(count))
    po.place_index(color, move)
    oo.place_index(color, move)
    sub, want_sub = tree.forward(move), om.forward(want_root, move)
    assert (sub is None) == (want_sub is None)
    if sub is not None:
        assert sub.total_count == want_sub.total_count
        sub.disqualify(361)                        # Player::predict_aux without allow_pass (self_play.rs:252-256)
        want_sub.disqualify(361)
        stub = hash_predictor(1.0)
        gv, gi, tree2, ge = pm.predict(pm.python_predictor(stub), po, 3 - color, starting_tree=sub, **kw)
        wv, wi, root2, we = om.predict(stub, oo, 3 - color, starting_tree=want_sub, **kw)
        count2, _, _ = tree2.children()
        assert (count2 == root2.count[:362]).all() and gi == wi and ge == we and count2[361] == 0


def test_tree_reuse_bit_exact():
    plays, komi = corpus_position(20, 60)
    po, oo = boards(plays, komi)
    color = po.to_move()
    stub = hash_predictor(1.5)
    kw = dict(deterministic=True, num_rollout=80, probes_per_round=2, leaf_symmetries=[4, 0, 1])
    gv, gi, tree, _ = pm.predict(pm.python_predictor(stub), po, color, **kw)
    wv, wi, root, _ = om.predict(stub, oo, color, **kw)
    assert gi == wi
    po.place_index(color, gi)
    oo.place_index(color, wi)
    sub, want_sub = tree.forward(gi), om.forward(root, wi)
    assert sub is not None and want_sub is not None and sub.total_count == want_sub.total_count
    gv, gi, tree2, ge = pm.predict(pm.python_predictor(stub), po, 3 - color, starting_tree=sub, **kw)
    wv, wi, root2, we = om.predict(stub, oo, 3 - color, starting_tree=want_sub, **kw)
    count, value, prior = tree2.children()
    assert (count == root2.count[:362]).all() and gi == wi and ge == we
    assert tree2.total_count == root2.total_count >= want_sub.total_count


def test_symmetric_position_folds_candidates():
    po, oo = boards([(BLACK, 9 * 19 + 9)])        # tengen: the full symmetry group survives
    tree, _ = assert_same_search(uniform_predictor(0.1), po, oo, WHITE, deterministic=True, num_rollout=60,
                                 leaf_symmetries=[0])
    _, _, prior = tree.children()
    assert np.isfinite(prior[:361]).sum() == 54   # 55 orbits minus the occupied centre


def test_degenerate_predictors():                 # lib.rs:245-281
    po, oo = boards([])
    v, i, tree, _ = pm.predict(pm.python_predictor(nan_predictor()), po, BLACK, deterministic=True, num_rollout=1600)
    assert v == 0.5 and i == 361 and tree.total_count == 0
    v, i, tree, _ = pm.predict(pm.python_predictor(fake_predictor(1, 0.6)), pgo.Board(0.5), BLACK, deterministic=True, num_rollout=40)
    assert i == 1


def test_visit_order_is_descending_prior_on_product():   # tree.rs:1773-1823 through the public search call
    """First round with many probes in flight: the leaves arrive in order of decreasing root prior."""
    seen = []

    def stub(feats):
        for f in feats:
            last = np.flatnonzero(f[:, 3].astype(np.float32))
            seen.append(int(last[0]) if len(last) else -1)
        return hash_predictor()(feats)

    po = pgo.Board(7.5)
    po.place(BLACK, 2, 3)                          # no symmetry
    _, _, tree, _ = pm.predict(pm.python_predictor(stub), po, WHITE, deterministic=True, num_rollout=2, probes_per_round=400,
                               leaf_symmetries=[0])
    _, _, prior = tree.children()
    first_round = [m for m in seen[8:] if m >= 0]
    assert len(first_round) > 100
    p = prior[first_round]
    assert (np.diff(p) <= 0).all()


def test_run_to_run_and_thread_count_determinism_of_self_play():
    stub = pm.python_predictor(hash_predictor())
    a, sgf_a = pm.self_play(stub, num_games=3, num_parallel=3, num_rollout=12, probes_per_round=2, max_plies=12, seed=7, num_threads=1)
    b, sgf_b = pm.self_play(stub, num_games=3, num_parallel=3, num_rollout=12, probes_per_round=2, max_plies=12, seed=7, num_threads=4)
    c, _ = pm.self_play(stub, num_games=3, num_parallel=3, num_rollout=12, probes_per_round=2, max_plies=12, seed=8, num_threads=4)
    assert a["digest"] == b["digest"] and sorted(sgf_a) == sorted(sgf_b)
    assert a["digest"] != c["digest"]
    # more games than slots, any number of groups and workers: whichever worker and group a game meets, it is the same game
    kw = dict(num_games=7, num_parallel=5, num_rollout=10, probes_per_round=2, max_plies=8, seed=7)
    ref, sgf_ref = pm.self_play(stub, num_threads=1, num_groups=1, **kw)
    for groups, threads in ((2, 3), (5, 8), (3, 2)):
        got, sgf_got = pm.self_play(stub, num_threads=threads, num_groups=groups, **kw)
        assert got["digest"] == ref["digest"] and sorted(sgf_got) == sorted(sgf_ref) and got["moves"] == ref["moves"]
        assert got["games_finished"] == 7
    assert a["games_finished"] == 3 and a["moves"] == 36
    assert all(g.startswith("(;GM[1]FF[4]SZ[19]RU[Chinese]KM[") and g.endswith(")") for g in sgf_a)


def test_gamma_sampler_statistics():
    """The product's own Dirichlet(0.03) sample (used when no noise is injected): normalised, sparse, seeded."""
    po = pgo.Board(7.5)
    po.place(WHITE, 2, 3)                          # no symmetry: the prior before the noise is uniform
    stub = pm.python_predictor(uniform_predictor())
    _, _, t1, _ = pm.predict(stub, po, BLACK, deterministic=False, num_rollout=1, seed=11, leaf_symmetries=[0])
    _, _, t2, _ = pm.predict(stub, po, BLACK, deterministic=False, num_rollout=1, seed=11, leaf_symmetries=[0])
    _, _, t3, _ = pm.predict(stub, po, BLACK, deterministic=False, num_rollout=1, seed=12, leaf_symmetries=[0])
    p1, p2, p3 = t1.children()[2], t2.children()[2], t3.children()[2]
    assert (p1 == p2).all() and (p1 != p3).any()
    f = np.isfinite(p1)
    assert abs(float(p1[f].sum()) - 1.0) < 1e-3
    base = 0.75 / f.sum()
    assert (p1[f] >= base * 0.999).all() and p1[f].max() > 0.05      # 25 % of the mass on a few points


# ---- transposition table (predictors/nn.rs:29-82, lru_cache.rs) ---------------------------------------------------------

def test_lru_kats_on_oracle_cache():       # lru_cache.rs:153-183
    c = om.Cache(1000)
    b = ogo.Board(7.5)
    class FakeBoard:
        def __init__(self, h): self.h = h
        def zobrist_hash(self): return self.h
    for i in range(20000):
        c.cache(FakeBoard(i), 1, 0, 0.0, np.zeros(362, np.float16))
    assert len(c.entries) == 1000
    c = om.Cache(10)
    for i in range(10): c.cache(FakeBoard(i), 1, 0, 0.0, np.zeros(362, np.float16))
    for i in range(2): c.fetch(FakeBoard(i), 1, 0)
    for i in range(6): c.cache(FakeBoard(i + 20), 1, 0, 0.0, np.zeros(362, np.float16))
    assert all(c.fetch(FakeBoard(i), 1, 0) is not None for i in (0, 1, 8, 9))
    assert all(c.fetch(FakeBoard(i), 1, 0) is None for i in range(2, 8))


@pytest.mark.parametrize("capacity", [200000, 48])
def test_search_with_transposition_table_bit_exact(capacity):
    """Two consecutive searches of one game share a table: root hits on the second search, leaf hits and (with the
    small capacity) evictions inside both -- visit counts, evaluated-position counts and hit counts are identical."""
    plays, komi = corpus_position(30, 50)
    po, oo = boards(plays, komi)
    color = po.to_move()
    stub = hash_predictor(1.2)
    pc, oc = pm.Cache(capacity), om.Cache(capacity)
    kw = dict(deterministic=True, num_rollout=110, probes_per_round=3, leaf_symmetries=[0, 4, 7, 2, 5])
    gv, gi, tree, ge = pm.predict(pm.python_predictor(stub), po, color, cache=pc, **kw)
    wv, wi, root, we = om.predict(stub, oo, color, cache=oc, **kw)
    count, _, _ = tree.children()
    assert (count == root.count[:362]).all() and gi == wi and ge == we
    assert pc.stats()["hits"] == oc.hits and pc.stats()["misses"] == oc.misses and pc.stats()["size"] == len(oc.entries)
    po.place_index(color, gi)
    oo.place_index(color, wi)
    sub, want_sub = tree.forward(gi), om.forward(root, wi)
    gv, gi, tree2, ge2 = pm.predict(pm.python_predictor(stub), po, 3 - color, starting_tree=sub, cache=pc, **kw)
    wv, wi, root2, we2 = om.predict(stub, oo, 3 - color, starting_tree=want_sub, cache=oc, **kw)
    count2, value2, prior2 = tree2.children()
    assert (count2 == root2.count[:362]).all() and gi == wi and ge2 == we2
    assert (prior2.view(np.uint32) == root2.prior[:362].view(np.uint32)).all()
    assert pc.stats()["hits"] == oc.hits > 0 and pc.stats()["size"] == len(oc.entries) <= capacity
    if capacity >= 200000:
        assert ge2 < ge                    # the second root (a leaf of the first search) came out of the table


def test_self_play_with_tables_is_reproducible_and_saves_evaluations():
    stub = pm.python_predictor(hash_predictor())
    kw = dict(num_games=2, num_parallel=2, num_rollout=16, probes_per_round=2, max_plies=10, seed=3)
    a, _ = pm.self_play(stub, num_threads=1, cache_capacity=4096, **kw)
    b, _ = pm.self_play(stub, num_threads=2, cache_capacity=4096, **kw)
    c, _ = pm.self_play(stub, num_threads=2, cache_capacity=0, **kw)
    assert a["digest"] == b["digest"] and a["cache_hits"] == b["cache_hits"] > 0
    assert a["evals"] < c["evals"] and c["cache_hits"] == 0


def test_striped_table_keeps_the_reference_semantics_per_stripe_and_survives_many_threads():
    """dg_cache_new_shared: a one-stripe table searches exactly like dg_cache_new (the reference's single LRU); a striped one
    shared by concurrent searches on several threads stays consistent (size <= capacity, hits + misses = lookups) and a
    second pass over the same positions is answered from it."""
    import threading
    plays, komi = corpus_position(30, 50)
    po, oo = boards(plays, komi)
    color = po.to_move()
    stub = hash_predictor(1.2)
    kw = dict(deterministic=True, num_rollout=110, probes_per_round=3, leaf_symmetries=[0, 4, 7, 2, 5])
    one, ref = pm.Cache(48, stripes=1), pm.Cache(48)
    _, i1, t1, e1 = pm.predict(pm.python_predictor(stub), po, color, cache=one, **kw)
    _, i2, t2, e2 = pm.predict(pm.python_predictor(stub), po, color, cache=ref, **kw)
    assert (t1.children()[0] == t2.children()[0]).all() and i1 == i2 and e1 == e2 and one.stats() == ref.stats()
    # many searches at once on one striped table (RandomPredictor: native, releases the GIL inside the search)
    shared = pm.Cache(20000, stripes=16)
    positions = []
    for game in (3, 8, 12, 30):
        pl, km = corpus_position(game, 40)
        positions.append(boards(pl, km)[0])
    evals = [[0, 0] for _ in positions]

    def worker(k, rnd):
        b = positions[k]
        _, _, _, ev = pm.predict(pm.RandomPredictor(), b, b.to_move(), deterministic=True, num_rollout=300, probes_per_round=4,
                                 leaf_symmetries=[k, 1, 5], cache=shared)
        evals[k][rnd] = ev

    for rnd in range(2):
        threads = [threading.Thread(target=worker, args=(k, rnd)) for k in range(len(positions))]
        [t.start() for t in threads]
        [t.join() for t in threads]
    st = shared.stats()
    assert 0 < st["size"] <= 20000 and st["hits"] > 0
    assert all(second < first for first, second in evals), evals      # the second pass finds the first one's evaluations


def test_self_play_with_one_shared_table():
    """cache_shared: one process-wide table for every game (predictors/nn.rs:48-50) instead of one per game."""
    kw = dict(num_games=12, num_parallel=2, num_rollout=24, probes_per_round=2, max_plies=8, seed=3, num_threads=2)
    own, _ = pm.self_play(pm.RandomPredictor(), cache_capacity=4096, **kw)
    shared, games = pm.self_play(pm.RandomPredictor(), cache_capacity=200000, cache_shared=16, **kw)
    none, _ = pm.self_play(pm.RandomPredictor(), cache_capacity=0, **kw)
    assert shared["games_finished"] == 12 and len(games) == 12 and shared["moves"] == none["moves"] == 96
    # every game opens on the same empty board: only a shared table lets the later games reuse the earlier games' opening
    # evaluations (the root's 8 symmetries and the first plies of the tree)
    assert shared["cache_hits"] > own["cache_hits"] > 0 and shared["evals"] < own["evals"] < none["evals"]


# ---- self-play records (self_play.rs:187-214, game_result.rs:23-43) --------------------------------------------------

def parse_record(sgf: str):
    import re
    from oracle import oracle as onn
    assert sgf.startswith("(;GM[1]FF[4]") and sgf.endswith(")")
    komi = float(re.search(r"KM\[([-0-9.]+)\]", sgf).group(1))
    moves = []
    for m in re.finditer(r";([BW])\[([a-s]{0,2})\]((?:[A-Z]+\[[^\]]*\])*)", sgf):
        color = 1 if m.group(1) == "B" else 2
        xy = m.group(2)
        index = 361 if not xy else (ord(xy[1]) - 97) * 19 + (ord(xy[0]) - 97)        # CGoban: x letter, y letter (sgf.rs:36-43)
        props = dict(re.findall(r"([A-Z]+)\[([^\]]*)\]", m.group(3)))
        moves.append((color, index, props))
    return komi, moves


@pytest.mark.parametrize("mode", ["search", "ex_it", "policy_only"])
def test_self_play_records_replay_legally_on_the_oracle(mode):
    from oracle import oracle as onn
    stub = pm.python_predictor(hash_predictor())
    kw = dict(num_games=2, num_parallel=2, probes_per_round=2, max_plies=14, seed=21, num_threads=2)
    if mode == "search":
        st, games = pm.self_play(stub, num_rollout=24, **kw)
    elif mode == "ex_it":
        st, games = pm.self_play(stub, num_rollout=24, ex_it=True, num_ex_it_rollout=40, **{**kw, "max_plies": 40})
    else:
        st, games = pm.self_play(stub, num_rollout=1, **kw)       # `--num-rollout 1`: play from the averaged policy
    assert st["games_finished"] == 2 and len(games) == 2
    for sgf in games:
        import re
        komi, moves = parse_record(sgf)
        assert len(moves) == (40 if mode == "ex_it" else 14)
        board = ogo.Board(komi)
        for ply, (color, index, props) in enumerate(moves):
            assert color == (1 if ply % 2 == 0 else 2)
            if index < 361:
                assert board.is_valid(color, index % 19, index // 19), (ply, index)
                board.place_index(color, index)
            assert "V" in props and -1.0 <= float(props["V"]) <= 1.0
            if mode != "policy_only":
                assert int(props["TV"]) >= 1
                dist = np.frombuffer(onn.b85_decode(props["P"].encode("ascii")), "<f2")[:362].astype(np.float32)   # as contrib/trainer reads it
                assert abs(float(dist.sum()) - 1.0) < 2e-2 and (dist >= 0).all()
                if index < 361:
                    assert dist[index] > 0
            else:
                assert "TV" not in props and "P" not in props
        # result and territory lists as game_result.rs:23-93 writes them, from the oracle's scoring of the final position
        assert re.search(r"RE\[([^\]]*)\]", sgf).group(1) == board.result()
        terr = board.territory()
        for tag, color in (("TB", 1), ("TW", 2)):
            m = re.search(tag + r"((?:\[[a-s]{2}\])+)", sgf)
            got = sorted((ord(p[1]) - 97) * 19 + ord(p[0]) - 97 for p in re.findall(r"\[([a-s]{2})\]", m.group(1))) if m else []
            assert got == np.flatnonzero(terr == color).tolist()
    if mode == "ex_it":
        assert st["searches"] > st["moves"]          # some positions were searched twice


# ---- whole games: product driver vs the oracle's self_play_one, same random stream --------------------------------------

def test_rng_restatement_matches_the_engine_stream():
    """oracle/rng.py reproduces csrc/search.h: Rng -- checked through its visible effects: komi draw and Dirichlet noise."""
    from oracle.rng import Rng
    po = pgo.Board(7.5)
    po.place(WHITE, 2, 3)
    oo = ogo.Board(7.5)
    oo.place(WHITE, 2, 3)
    stub = uniform_predictor()
    _, _, tree, _ = pm.predict(pm.python_predictor(stub), po, BLACK, deterministic=False, num_rollout=1, seed=77)
    _, policy, _ = om.full_forward(stub, 0, oo, BLACK)
    eta = Rng(77).dirichlet(policy, float(np.float32(0.03)))
    om.dirichlet_mix(policy[:362], eta, 0.25)
    assert (tree.children()[2].view(np.uint32) == policy[:362].view(np.uint32)).all()


@pytest.mark.parametrize("probes,rollouts,cache", [(1, 20, 0), (3, 20, 0), (2, 20, 512), (2, 1, 0)])
def test_whole_game_matches_the_oracle_game(probes, rollouts, cache):
    """One self-play game, move for move: komi, Dirichlet noise, leaf symmetries, stochastic opening moves, tree re-use,
    scoring search with pass disqualified, the per-move rollout budget -- everything drawn from the same seeded stream."""
    from oracle.rng import Rng
    stub = hash_predictor()
    seed, plies = 5, 12
    st, games = pm.self_play(pm.python_predictor(stub), num_games=1, num_parallel=1, num_rollout=rollouts, probes_per_round=probes,
                             max_plies=plies, seed=seed, num_threads=1, cache_capacity=cache)
    komi, moves = parse_record(games[0])
    game_rng = Rng((seed * 0x9e3779b97f4a7c15 + 0 * 0xd1342543de82ef95 + 1) & ((1 << 64) - 1))     # Driver::start_game, game id 0
    want_komi, want_moves = om.self_play_one(stub, game_rng, num_rollout=rollouts, probes_per_round=probes, max_plies=plies,
                                             cache=om.Cache(cache) if cache else None)
    assert komi == want_komi
    assert [(c, i) for c, i, _ in moves] == want_moves


@pytest.mark.parametrize("rollouts,probes,plies", [(24, 2, 60), (1, 2, 120)])
def test_whole_game_with_ex_it_matches_the_oracle_game(rollouts, probes, plies):
    """`--ex-it` (BASELINE configs[4]; rollouts = 1 is its policy-only play): which moves get the second, deeper search (one f32
    draw per eligible move), the seed it runs under, and what the record then holds -- TV[] = the rollouts of the recorded tree
    and P[] = its visit distribution -- while the game goes on with the first search's move (self_play.rs:287-319, 341-346,
    384-389).  Move for move against the oracle's self_play_one on the same random stream."""
    from oracle import oracle as onn
    from oracle.rng import Rng
    stub = hash_predictor()
    seed = 9
    st, games = pm.self_play(pm.python_predictor(stub), num_games=1, num_parallel=1, num_rollout=rollouts, probes_per_round=probes,
                             max_plies=plies, seed=seed, num_threads=1, ex_it=True, num_ex_it_rollout=40)
    komi, moves = parse_record(games[0])
    recorded = []
    want_komi, want_moves = om.self_play_one(stub, Rng((seed * 0x9e3779b97f4a7c15 + 1) & ((1 << 64) - 1)), num_rollout=rollouts,
                                             probes_per_round=probes, max_plies=plies, ex_it=True, num_ex_it_rollout=40, recorded=recorded)
    assert komi == want_komi
    assert [(c, i) for c, i, _ in moves] == want_moves
    expanded = 0
    for (_, _, props), rec in zip(moves, recorded):
        if rec is None or rec[0] <= 1:
            assert "TV" not in props and "P" not in props
            continue
        assert int(props["TV"]) == rec[0]
        dist = np.frombuffer(onn.b85_decode(props["P"].encode("ascii")), "<f2")[:362]
        assert (dist.view(np.uint16) == rec[1].astype(np.float16).view(np.uint16)).all()
        expanded += 1
    assert st["searches"] > len(moves) if rollouts > 1 else st["searches"] > 0     # some positions were searched a second time
    assert expanded >= (len(moves) if rollouts > 1 else 1)


# ---- the games themselves, pinned (tests/golden/selfplay_digests.json, tools/make_selfplay_digests.py) ---------------

def test_self_play_digests_match_the_golden_file():
    """Whole self-play runs with the RandomPredictor give the games they gave when the fixture was written -- for every
    number of worker threads and groups."""
    import json
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "selfplay_digests.json")
    for k, case in enumerate(json.load(open(path))):
        st, games = pm.self_play(pm.RandomPredictor(), num_threads=1 + k % 3, num_groups=1 + k % 4, **case["config"])
        assert f"{st['digest']:016x}" == case["digest"], case["config"]
        assert int(st["moves"]) == case["moves"] and len(games) == case["games"]
