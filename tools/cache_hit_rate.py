#!/usr/bin/env python
"""Hit rate of the process-wide transposition table (`NnPredictor`'s 200,000-entry LRU, predictors/nn.rs:29-82) on
positions where transpositions exist: the consecutive positions of the fixture games (dg_tests/fixtures/example_games.sgf),
each searched with `--num-rollout` rollouts without tree re-use, all searches sharing ONE striped table -- what a match or
an analysis session does.  Host only (RandomPredictor); the table's behaviour does not depend on the evaluator.

    python tools/cache_hit_rate.py [--games 12] [--plies 80] [--rollouts 800] [--threads 4] [--capacity 200000]
Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=12)
    ap.add_argument("--plies", type=int, default=80)
    ap.add_argument("--rollouts", type=int, default=800)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--capacity", type=int, default=200000)
    ap.add_argument("--stripes", type=int, default=64)
    ap.add_argument("--sharpness", type=float, default=0.0, help="> 0: PeakedPredictor(sharpness) instead of the RandomPredictor")
    args = ap.parse_args()
    from dream_go_b200 import go as pgo, mcts
    from oracle import go as ogo
    games = ogo.load_games()[:args.games]
    table = mcts.Cache(args.capacity, stripes=args.stripes)
    predictor = mcts.PeakedPredictor(args.sharpness) if args.sharpness > 0 else mcts.RandomPredictor()
    evals = [0] * len(games)
    searches = [0] * len(games)

    def play(k):
        colors, moves, komi = games[k]
        board = pgo.Board(komi)
        for c, m in list(zip(colors, moves))[:args.plies]:
            _, _, _, ev = mcts.predict(predictor, board, int(c), deterministic=True, num_rollout=args.rollouts,
                                       probes_per_round=8, seed=k + 1, cache=table)
            evals[k] += ev
            searches[k] += 1
            if m < 361:
                board.place_index(int(c), int(m))

    t0 = time.perf_counter()
    pending = list(range(len(games)))
    lock = threading.Lock()

    def worker():
        while True:
            with lock:
                if not pending:
                    return
                k = pending.pop()
            play(k)

    threads = [threading.Thread(target=worker) for _ in range(args.threads)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    st = table.stats()
    lookups = st["hits"] + st["misses"]
    print(json.dumps({"predictor": f"PeakedPredictor({args.sharpness})" if args.sharpness > 0 else "RandomPredictor (near-uniform policy)",
                      "workload": f"{sum(searches)} searches of {args.rollouts} rollouts over the first {args.plies} plies of {len(games)} fixture games, "
                                  f"no tree re-use, one table of {args.capacity} entries in {args.stripes} stripes shared by {args.threads} threads",
                      "lookups": lookups, "hits": st["hits"], "hit_rate": st["hits"] / max(lookups, 1), "entries": st["size"],
                      "network_evaluations": sum(evals), "evaluations_saved_frac": st["hits"] / max(st["hits"] + sum(evals), 1),
                      "seconds": time.perf_counter() - t0}))


if __name__ == "__main__":
    main()
