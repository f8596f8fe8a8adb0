"""Pins the tree-search restatement (oracle/mcts.py) against the invariants the reference's own tests hold
(src/libdg_mcts/tree.rs:1758-1946, lib.rs:245-281, predictor.rs:99-107, choose.rs tests).  CPU only."""
import numpy as np
import pytest

from oracle import go, mcts
from mcts_common import fake_predictor, nan_predictor, hash_predictor

BLACK, WHITE = 1, 2
F = np.float32


def prior_distribution(seed):            # tree.rs:1758-1771
    rng = np.random.default_rng(seed)
    prior = rng.random(362).astype(F)
    return (prior / prior.sum()).astype(F)


def test_visit_order_is_descending_prior():   # tree.rs:1773-1823
    root = mcts.Node(BLACK, 0.5, prior_distribution(1))
    choices = []
    while True:
        status, trace = mcts.probe(root, go.Board(7.5))
        if status != "found":
            break
        assert len(trace) == 1
        i = trace[0][1]
        assert i not in choices
        choices.append(i)
        assert root.vcount[i] == mcts.VLOSS_CNT
        assert root.vtotal_count == len(choices) * mcts.VLOSS_CNT
        assert all(root.prior[o] >= root.prior[i] for o in choices)
        assert len(choices) <= 362
    assert len(choices) > 300          # ends at the first conflict (an in-flight child outscoring the small priors)
    assert all(root.vcount[o] == mcts.VLOSS_CNT for o in choices)


def test_virtual_loss():                      # tree.rs:1830-1862
    board = go.Board(7.5)
    root = mcts.Node(BLACK, 0.5, prior_distribution(1))
    status, trace = mcts.probe(root, board)
    assert status == "found"
    i = trace[0][1]
    assert root.vcount[i] == 32 and root.vtotal_count == 32
    mcts.insert(trace, BLACK, 0.9, prior_distribution(2))
    assert root.vcount[i] == 0 and root.vtotal_count == 0 and root.count[i] == 1 and root.total_count == 1


def test_value_update():                      # tree.rs:1869-1925
    board = go.Board(7.5)
    prior = np.zeros(362, F)
    prior[60] = 1.0
    root = mcts.Node(BLACK, 0.5, prior)
    other = np.zeros(362, F)
    other[61] = other[62] = 0.5
    _, trace = mcts.probe(root, board)
    mcts.insert(trace, BLACK, 0.9, other)
    assert 0.8999 <= root.value[60] <= 0.9001 and root.count[60] == 1 and root.total_count == 1
    assert root.vcount[60] == 0 and root.vtotal_count == 0
    _, t1 = mcts.probe(root, go.Board(7.5))
    _, t2 = mcts.probe(root, go.Board(7.5))
    assert t1[0][1] == 60 and t2[0][1] == 60 and {t1[1][1], t2[1][1]} == {61, 62}
    assert root.value[60] == F(0.9) and root.count[60] == 1 and root.total_count == 1
    assert root.vcount[60] == 64 and root.vtotal_count == 64
    mcts.insert(t1, WHITE, 0.2, other)
    assert root.value[60] == F(0.85) and root.count[60] == 2 and root.vcount[60] == 32 and root.vtotal_count == 32
    mcts.insert(t2, WHITE, 0.3, other)
    assert root.value[60] == F(0.8) and root.count[60] == 3 and root.total_count == 3
    assert root.vcount[60] == 0 and root.vtotal_count == 0


def test_undo_trace():                        # tree.rs:1927-1946
    board = go.Board(7.5)
    prior = np.zeros(362, F)
    prior[60] = 1.0
    root = mcts.Node(BLACK, 0.5, prior)
    assert mcts.probe(root, board)[0] == "found"
    assert mcts.probe(root, board)[0] != "found"
    assert root.vtotal_count == mcts.VLOSS_CNT


def test_no_allowed_moves():                  # lib.rs:245-262
    root = mcts.Node(BLACK, 0.0, np.ones(362, F))
    for i in range(362):
        root.disqualify(i)
    _, _, tree, _ = mcts.predict(hash_predictor(), go.Board(7.5), BLACK, deterministic=True, num_rollout=100, starting_tree=root)
    value, index = mcts.best(tree, 0.0, 0.0)
    assert value == -np.inf and index == 361


def test_no_finite_candidates():              # lib.rs:264-281
    value, index, root, _ = mcts.predict(nan_predictor(), go.Board(7.5), BLACK, deterministic=True, num_rollout=1600)
    assert value == 0.5 and index == 361 and root.total_count == 0 and root.vtotal_count == 0


def test_fake_predictor_search_plays_the_predicted_point():   # self_play.rs:558-591 (played_from_mcts set-up)
    value, index, root, _ = mcts.predict(fake_predictor(1, 0.6), go.Board(0.5), BLACK, deterministic=True, num_rollout=40)
    assert index == 1 and 0.0 < float(value) < 1.0 and root.total_count >= 1


def test_choose_respects_cutoff_and_temperature():   # choose.rs tests
    items = [0.0, 10.0, 20.0, 70.0]
    assert mcts.choose(items, 0.5, 1.0, 0.0) == 3
    assert mcts.choose(items, 0.0, 1.0, 0.05) == 1
    assert mcts.choose(items, 0.0, 1.0, 0.99) == 3
    assert mcts.choose([0.0, 0.0], 0.5, 1.0, 0.5) is None      # 0/0 -> NaN never reaches `at`


def test_interpolation_schedule():            # config.rs:297-312
    assert mcts.get_intp_value(mcts.UCT_EXP, 0) == F(0.77392)
    assert mcts.get_intp_value(mcts.UCT_EXP, 800) == F(1.05439)
    assert abs(float(mcts.get_intp_value(mcts.UCT_EXP, 400)) - 0.5 * (0.77392 + 1.05439)) < 1e-6
    assert mcts.get_intp_value(mcts.UCT_EXP, 10 ** 6) == F(0.764326)
