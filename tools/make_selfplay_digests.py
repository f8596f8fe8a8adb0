#!/usr/bin/env python
"""Pins the games the host half plays: order-independent digests of whole self-play runs with the RandomPredictor
(`dg_random_predict`, no device) for a few configurations -> tests/golden/selfplay_digests.json.  The digest hashes every
finished game's move list, so any change of the board rules, the feature-independent parts of the search, the prior
construction, the random streams or the record of a game shows up here.  Regenerate ONLY when such a change is intended
(the product-vs-oracle tests say whether it is right; this file says whether it is the same).

    python tools/make_selfplay_digests.py            # rewrites the fixture
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    dict(num_games=6, num_parallel=4, num_rollout=40, probes_per_round=4, max_plies=30, seed=11),
    dict(num_games=4, num_parallel=4, num_rollout=120, probes_per_round=8, max_plies=16, seed=12),
    dict(num_games=5, num_parallel=3, num_rollout=24, probes_per_round=2, max_plies=60, seed=13, ex_it=True, num_ex_it_rollout=48),
    dict(num_games=8, num_parallel=8, num_rollout=1, probes_per_round=1, max_plies=120, seed=14),
    dict(num_games=3, num_parallel=3, num_rollout=30, probes_per_round=3, max_plies=40, seed=15, cache_capacity=256),
]


def run(case):
    from dream_go_b200 import mcts
    st, games = mcts.self_play(mcts.RandomPredictor(), num_threads=2, **case)
    return {"digest": f"{st['digest']:016x}", "moves": int(st["moves"]), "games": len(games)}


def main():
    out = [{"config": c, **run(c)} for c in CASES]
    path = os.path.join(ROOT, "tests", "golden", "selfplay_digests.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print(path, [o["digest"] for o in out])


if __name__ == "__main__":
    main()
