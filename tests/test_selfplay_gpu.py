"""The whole hot path on the device: real feature planes -> engine -> priors -> tree search.

* a search driven by the B200 engine gives the SAME visit counts as the oracle search fed by the same engine
  ("MCTS visit counts bit-exact under a fixed seed");
* self-play on the engine is reproducible run to run and independent of the host thread count (no float atomics
  anywhere on the device, every game owns its random stream)."""
import numpy as np
import pytest

from dream_go_b200 import go as pgo, mcts as pm, nn, weights
from oracle import go as ogo, mcts as om, oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(small_net):
    net = nn.Network.from_tensors(small_net, max_batch=256, num_workspaces=2)
    yield net
    net.close()


def test_real_feature_planes_match_oracle_network(engine, small_net):
    """Real positions (not Bernoulli noise) through pack_compact_kernel + tower + heads vs the CPU oracle network."""
    colors, moves, komi = ogo.load_games()[2]
    out = pgo.replay(colors[:64], moves[:64], komi, features=True)
    want_feats = ogo.replay(colors[:64], moves[:64], komi, features=True)["features"]
    got = engine.forward_packed(out["features"])
    want_v, want_p = oracle.OracleNetwork(small_net).forward(want_feats)
    v, p = got.value.astype(np.float32), got.policy.reshape(-1, 362).astype(np.float32)
    assert np.abs(v - want_v.astype(np.float32)).max() <= 4e-3
    assert (np.abs(p - want_p.astype(np.float32)) <= 1e-3 + 1e-2 * want_p.astype(np.float32)).all()


@pytest.mark.parametrize("probes,rollouts", [(1, 150), (4, 150), (8, 800)])
def test_search_on_engine_matches_oracle_search(engine, probes, rollouts):
    """(8, 800) is the search every BASELINE config runs: n passes the first knot of the UCT_EXP / FPU_REDUCE tables
    (src/libdg_utils/config.rs:181-195, 297-336)."""
    ogo.use_default_zobrist()
    colors, moves, komi = ogo.load_games()[9]
    po, oo = pgo.Board(komi), ogo.Board(komi)
    for c, m in zip(colors[:40], moves[:40]):
        if m < 361:
            po.place_index(int(c), int(m))
            oo.place_index(int(c), int(m))
    color = po.to_move()

    def engine_on_features(feats):          # the oracle search evaluates fp16 feature tensors on the same engine
        with engine.get_workspace(len(feats)) as ws:
            value, policy = nn.forward(ws, np.ascontiguousarray(feats)).unwrap()
        return value, policy.reshape(-1, 362)

    kw = dict(deterministic=True, num_rollout=rollouts, probes_per_round=probes, leaf_symmetries=[0, 3, 6, 1, 5])
    want_v, want_i, want_root, want_evals = om.predict(engine_on_features, oo, color, **kw)
    got_v, got_i, tree, got_evals = pm.predict(pm.EnginePredictor(engine), po, color, **kw)
    count, value, prior = tree.children()
    assert (count == want_root.count[:362]).all()
    assert got_i == want_i and got_evals == want_evals
    assert np.float32(got_v).view(np.uint32) == np.float32(want_v).view(np.uint32)
    assert (prior.view(np.uint32) == want_root.prior[:362].view(np.uint32)).all()


def test_self_play_on_engine_is_reproducible(engine):
    kw = dict(num_games=4, num_parallel=4, num_rollout=40, probes_per_round=4, max_plies=16, seed=5)
    a, sgf_a = pm.self_play(pm.EnginePredictor(engine), num_threads=1, **kw)
    b, sgf_b = pm.self_play(pm.EnginePredictor(engine), num_threads=4, **kw)
    assert a["digest"] == b["digest"] and sorted(sgf_a) == sorted(sgf_b)
    assert a["games_finished"] == 4 and a["moves"] == 64 and a["evals"] == b["evals"] > 0


def test_device_features_give_the_same_search_and_games(engine):
    """Raw positions (planes + legal moves derived on the device) vs host feature planes: identical trees, identical games."""
    colors, moves, komi = ogo.load_games()[13]
    po = pgo.Board(komi)
    for c, m in zip(colors[:70], moves[:70]):
        if m < 361:
            po.place_index(int(c), int(m))
    color = po.to_move()
    for kw in (dict(deterministic=True, num_rollout=120, probes_per_round=4, leaf_symmetries=[2, 0, 7, 5]),
               dict(search=1, deterministic=False, num_rollout=80, probes_per_round=2, seed=9)):
        v1, i1, t1, e1 = pm.predict(pm.EnginePredictor(engine), po, color, **kw)
        c1, val1, p1 = t1.children()
        for predictor in (pm.EngineRawPredictor(engine), pm.EnginePriorPredictor(engine)):     # planes / planes + priors on the device
            v2, i2, t2, e2 = pm.predict(predictor, po, color, **kw)
            c2, val2, p2 = t2.children()
            assert (c1 == c2).all() and i1 == i2 and e1 == e2 and v1 == v2
            assert (p1.view(np.uint32) == p2.view(np.uint32)).all() and (val1.view(np.uint32) == val2.view(np.uint32)).all()
    kw = dict(num_games=4, num_parallel=4, num_rollout=40, probes_per_round=4, max_plies=24, seed=11, num_threads=4)
    a, sgf_a = pm.self_play(pm.EnginePredictor(engine), **kw)
    b, sgf_b = pm.self_play(pm.EngineRawPredictor(engine), **kw)
    c, sgf_c = pm.self_play(pm.EnginePriorPredictor(engine), **kw)
    assert a["digest"] == b["digest"] == c["digest"] and sorted(sgf_a) == sorted(sgf_b) == sorted(sgf_c)
    assert a["evals"] == b["evals"] == c["evals"]


@pytest.mark.parametrize("flags", [0, nn.FLAG_NO_GRAPH | nn.FLAG_BLOCKING_SYNC])
def test_leaf_batch_queue_driver_plays_the_same_games(small_net, flags):
    """The product path (dg_selfplay_run_engine: leaf-batch queue, graph launches, completion flags, no device threads)
    plays the games of the blocking-call drivers, with host or device priors, for any number of workers and groups."""
    net = nn.Network.from_tensors(small_net, max_batch=128, num_workspaces=4, flags=flags)
    try:
        kw = dict(num_games=6, num_parallel=5, num_rollout=40, probes_per_round=4, max_plies=20, seed=11)
        ref, sgf_ref = pm.self_play(pm.EnginePredictor(net), num_threads=2, **kw)
        # (None = DG_SELFPLAY_AUTO_PRIORS: the driver moves the prior construction between host and device as it sees fit)
        for priors, ladders, threads, groups in ((False, False, 1, 0), (True, False, 4, 2), (False, True, 3, 4), (True, True, 2, 3),
                                                 (None, False, 2, 4), (None, None, 8, 2)):
            got, sgf = pm.self_play(pm.EngineQueue(net, device_priors=priors, device_ladders=ladders), num_threads=threads,
                                    num_groups=groups, **kw)
            assert got["digest"] == ref["digest"] and sorted(sgf) == sorted(sgf_ref), (priors, ladders, threads, groups)
            assert got["evals"] == ref["evals"] and got["games_finished"] == 6 and got["moves"] == ref["moves"]
        # a deadline in the middle of the run: the batches in flight are dropped, nothing hangs
        cut, _ = pm.self_play(pm.EngineQueue(net), num_threads=2, max_seconds=0.3, **{**kw, "num_games": 1000, "max_plies": 722})
        assert 0 < cut["evals"] and cut["seconds"] < 5.0
    finally:
        net.close()


def test_self_play_on_two_engines_of_one_process(small_net):
    """`NnPredictor` drives every GPU from one process, round-robin (predictors/nn.rs:84-92): dg_selfplay_run_engine with two
    engines on two devices (groups dealt round-robin, host threads shared) plays the games one engine plays."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nets = [nn.Network.from_tensors(small_net, device=d, max_batch=128, num_workspaces=4) for d in (0, 1)]
    try:
        kw = dict(num_games=8, num_parallel=6, num_rollout=40, probes_per_round=4, max_plies=20, seed=13, num_threads=4)
        one, sgf_one = pm.self_play(pm.EngineQueue(nets[0], device_priors=False), **kw)
        for priors in (False, True):
            two, sgf_two = pm.self_play(pm.EngineQueue(nets, device_priors=priors), **kw)
            assert two["digest"] == one["digest"] and sorted(sgf_two) == sorted(sgf_one) and two["evals"] == one["evals"]
        # and both devices really worked: a deadline run with many games keeps both engines' batches in flight
        big, _ = pm.self_play(pm.EngineQueue(nets), max_seconds=1.0, **{**kw, "num_games": 10000, "num_parallel": 64, "max_plies": 722})
        assert big["evals"] > 0 and big["rounds"] > 4
    finally:
        for net in nets:
            net.close()


def test_whole_game_on_engine_matches_oracle_game(engine):
    """A self-play game on the engine, move for move against the oracle's self_play_one fed by the same engine."""
    import re
    from oracle.rng import Rng
    ogo.use_default_zobrist()

    def engine_on_features(feats):
        with engine.get_workspace(len(feats)) as ws:
            value, policy = nn.forward(ws, np.ascontiguousarray(feats)).unwrap()
        return value, policy.reshape(-1, 362)

    seed, plies, rollouts = 3, 10, 48
    kw = dict(num_games=1, num_parallel=1, num_rollout=rollouts, probes_per_round=2, max_plies=plies, seed=seed, num_threads=1)
    _, games = pm.self_play(pm.EngineRawPredictor(engine), **kw)
    _, games_packed = pm.self_play(pm.EnginePredictor(engine), **kw)
    assert games == games_packed
    komi = float(re.search(r"KM\[([-0-9.]+)\]", games[0]).group(1))
    moves = [(1 if m.group(1) == "B" else 2, 361 if not m.group(2) else (ord(m.group(2)[1]) - 97) * 19 + ord(m.group(2)[0]) - 97)
             for m in re.finditer(r";([BW])\[([a-s]{0,2})\]", games[0])]
    game_rng = Rng((seed * 0x9e3779b97f4a7c15 + 1) & ((1 << 64) - 1))
    want_komi, want_moves = om.self_play_one(engine_on_features, game_rng, num_rollout=rollouts, probes_per_round=2, max_plies=plies)
    assert komi == want_komi and moves == want_moves


def test_self_play_on_engine_with_one_shared_table(small_net):
    """The process-wide transposition table (`cache_shared`, predictors/nn.rs:29-82) on the engine path: every game is played
    to the end with legal moves, later games reuse the evaluations of earlier ones (all start from the empty board), fewer
    positions reach the device than without a table."""
    import re
    net = nn.Network.from_tensors(small_net, max_batch=128, num_workspaces=4)
    try:
        kw = dict(num_games=8, num_parallel=2, num_rollout=24, probes_per_round=2, max_plies=8, seed=3, num_threads=2)
        plain, _ = pm.self_play(pm.EngineQueue(net, device_priors=False), **kw)
        shared, games = pm.self_play(pm.EngineQueue(net, device_priors=False), cache_capacity=200000, cache_shared=16, **kw)
        assert shared["games_finished"] == 8 and len(games) == 8 and shared["moves"] == plain["moves"] == 64
        assert shared["cache_hits"] > 0 and shared["evals"] < plain["evals"]
        for sgf in games:
            komi = float(re.search(r"KM\[([-0-9.]+)\]", sgf).group(1))
            board = ogo.Board(komi)
            for m in re.finditer(r";([BW])\[([a-s]{0,2})\]", sgf):
                color = 1 if m.group(1) == "B" else 2
                if m.group(2):
                    x, y = ord(m.group(2)[0]) - 97, ord(m.group(2)[1]) - 97
                    assert board.is_valid(color, x, y)
                    board.place(color, x, y)
    finally:
        net.close()
