#!/usr/bin/env python
"""Soak run of the whole-game parity check (tests/test_mcts_parity.py::test_whole_game_matches_the_oracle_game) at sizes the
test suite cannot afford: one self-play game of the product driver (csrc/search_api.cpp) against the oracle's self_play_one
(oracle/mcts.py), move for move, with the BASELINE rollout budget, long games, every probes-per-round setting and with a
transposition table.  CPU only (stub predictor: a deterministic function of the fp16 feature tensor).

    python tools/soak_parity.py [--out profiles/r02_soak_whole_games_vs_oracle.log] [--quick]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mcts_common import hash_predictor                      # noqa: E402
from test_mcts_parity import parse_record                   # noqa: E402
from dream_go_b200 import mcts as pm                        # noqa: E402
from oracle import mcts as om                               # noqa: E402
from oracle.rng import Rng                                  # noqa: E402

# (seed, plies, rollouts, probes per round, table capacity)
CASES = [
    (11, 60, 800, 8, 0),          # the BASELINE budget with the default probes per round
    (12, 60, 800, 1, 0),          # the sequential algorithm
    (13, 120, 200, 4, 0),         # into the middle game: ladders, captures, ko
    (14, 120, 200, 8, 4096),      # with a transposition table (hits, evictions, tree re-use)
    (15, 300, 60, 3, 0),          # long: scoring search, passes
    (16, 300, 60, 8, 512),
    (17, 722, 24, 2, 0),          # to the stone cap
    (18, 40, 1700, 8, 0),         # past the second knot of the schedules
]
QUICK = [(21, 30, 100, 4, 0), (22, 30, 100, 8, 256)]
# --ex-it (BASELINE configs[4]): (seed, plies, rollouts, probes per round, rollouts of the second search)
EX_IT = [(31, 160, 100, 8, 300), (32, 722, 1, 8, 200), (33, 250, 40, 4, 800)]


def run_case(seed, plies, rollouts, probes, cache):
    stub = hash_predictor()
    t0 = time.time()
    st, games = pm.self_play(pm.python_predictor(stub), num_games=1, num_parallel=1, num_rollout=rollouts, probes_per_round=probes,
                             max_plies=plies, seed=seed, num_threads=1, cache_capacity=cache)
    t1 = time.time()
    komi, moves = parse_record(games[0])
    game_rng = Rng((seed * 0x9e3779b97f4a7c15 + 1) & ((1 << 64) - 1))          # Driver::start_game, game id 0
    want_komi, want_moves = om.self_play_one(stub, game_rng, num_rollout=rollouts, probes_per_round=probes, max_plies=plies,
                                             cache=om.Cache(cache) if cache else None)
    t2 = time.time()
    got = [(c, i) for c, i, _ in moves]
    ok = komi == want_komi and got == want_moves
    first = next((k for k, (a, b) in enumerate(zip(got, want_moves)) if a != b), None)
    return ok, (f"seed {seed} plies<= {plies} rollouts {rollouts} probes {probes} table {cache}: {len(got)} moves, {int(st['evals'])} evaluations, "
                f"{int(st['cache_hits'])} table hits, komi {komi}; product {t1 - t0:.1f} s, oracle {t2 - t1:.1f} s -> "
                + ("IDENTICAL" if ok else f"DIFFERENT (first difference at move {first}, lengths {len(got)} / {len(want_moves)})"))


def run_ex_it_case(seed, plies, rollouts, probes, deep):
    import numpy as np
    from oracle import oracle as onn
    stub = hash_predictor()
    t0 = time.time()
    st, games = pm.self_play(pm.python_predictor(stub), num_games=1, num_parallel=1, num_rollout=rollouts, probes_per_round=probes,
                             max_plies=plies, seed=seed, num_threads=1, ex_it=True, num_ex_it_rollout=deep)
    t1 = time.time()
    komi, moves = parse_record(games[0])
    recorded = []
    want_komi, want_moves = om.self_play_one(stub, Rng((seed * 0x9e3779b97f4a7c15 + 1) & ((1 << 64) - 1)), num_rollout=rollouts,
                                             probes_per_round=probes, max_plies=plies, ex_it=True, num_ex_it_rollout=deep, recorded=recorded)
    t2 = time.time()
    got = [(c, i) for c, i, _ in moves]
    ok = komi == want_komi and got == want_moves and len(recorded) == len(moves)
    records = 0
    if ok:
        for (_, _, props), rec in zip(moves, recorded):
            if rec is None or rec[0] <= 1:
                ok = ok and "TV" not in props and "P" not in props
                continue
            dist = np.frombuffer(onn.b85_decode(props["P"].encode("ascii")), "<f2")[:362]
            ok = ok and int(props["TV"]) == rec[0] and bool((dist.view(np.uint16) == rec[1].astype(np.float16).view(np.uint16)).all())
            records += 1
    return ok, (f"--ex-it seed {seed} plies<= {plies} rollouts {rollouts} probes {probes} second search {deep}: {len(got)} moves, "
                f"{int(st['searches'])} searches ({int(st['searches']) - len(got)} second searches), "
                f"{records} records with TV[] / P[] compared, {int(st['evals'])} evaluations; product {t1 - t0:.1f} s, oracle {t2 - t1:.1f} s -> "
                + ("IDENTICAL" if ok else "DIFFERENT"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--ex-it", action="store_true", help="the --ex-it cases only (appended to --out)")
    args = ap.parse_args()
    out = open(args.out, "a" if args.ex_it else "w") if args.out else None
    bad = 0
    for case in (EX_IT if args.ex_it else QUICK if args.quick else CASES):
        ok, line = (run_ex_it_case if args.ex_it else run_case)(*case)
        bad += not ok
        print(line, flush=True)
        if out:
            out.write(line + "\n")
            out.flush()
    tail = f"{'ALL IDENTICAL' if not bad else str(bad) + ' DIFFERENT'}"
    print(tail)
    if out:
        out.write(tail + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
