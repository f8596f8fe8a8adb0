#!/usr/bin/env python
"""Randomised product-vs-oracle search parity (tests/test_mcts_parity.py::assert_same_search on random inputs): positions from
the fixture games at random plies and from random playouts, either search kind, with and without noise / temperature, rollout
budgets 1..500, 1..12 probes per round, random leaf symmetries, sharp / flat / tied stub predictors.  For every case the visit
count of every root child, the priors, the values of the visited children, the chosen move, its value and the number of
evaluated positions must be bit-identical.  CPU only.

    python tools/fuzz_search_parity.py [--cases 150] [--seed 1] [--out profiles/r02_fuzz_search_parity.log]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_mcts_parity as T                                # noqa: E402
import test_go_parity as G                                  # noqa: E402
from mcts_common import dirichlet_sample, hash_predictor    # noqa: E402
from oracle import go as ogo, mcts as om                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=150)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    ogo.use_default_zobrist()
    rng = np.random.default_rng(args.seed)
    games = ogo.load_games()
    lines, bad, evals_total = [], 0, 0
    t0 = time.time()
    for k in range(args.cases):
        source = rng.random()
        if source < 0.12:
            # a board with a symmetry (candidates folded onto orbit representatives, policy_helper.rs:54-72): whole orbits of
            # random points under a random subgroup of the eight transforms
            group = [[0, 1], [0, 2], [0, 3], [0, 4], [0, 6], [0, 1, 2, 6], [0, 3, 4, 6], list(range(8))][int(rng.integers(0, 8))]
            plays, komi, taken = [], 7.5, set()
            for _ in range(int(rng.integers(0, 14))):
                p0, c = int(rng.integers(0, 361)), int(rng.integers(1, 3))
                orbit = sorted({ogo.symmetry_apply(t, p0) for t in group})
                if not taken & set(orbit):
                    taken |= set(orbit)
                    plays += [(c, q) for q in orbit]
            where = f"symmetric under {group}"
        elif source < 0.75:
            g = int(rng.integers(0, len(games)))
            plies = int(rng.integers(0, len(games[g][1])))
            plays, komi = T.corpus_position(g, plies)
            where = f"game {g} ply {plies}"
        else:
            s, plies = int(rng.integers(0, 1 << 20)), int(rng.integers(0, 320))
            colors, moves = G.random_playout(s, plies, pass_rate=0.0)
            plays, komi = [(c, m) for c, m in zip(colors, moves) if m < 361], 7.5
            where = f"playout {s} ply {plies}"
        po, oo = T.boards(plays, komi)
        color = po.to_move()
        search = int(rng.integers(0, 2))
        deterministic = bool(rng.random() < 0.5)
        kind = rng.choice(["sharp", "flat", "tied"])
        stub = hash_predictor(4.0) if kind == "sharp" else hash_predictor(0.5) if kind == "flat" else T.tied_predictor()
        kw = dict(search=search, deterministic=deterministic, num_rollout=int(rng.choice([1, 2, 7, 30, 90, 200, 500])),
                  probes_per_round=int(rng.choice([1, 2, 3, 4, 8, 12])),
                  leaf_symmetries=[int(x) for x in rng.integers(0, 8, size=int(rng.integers(1, 9)))])
        if not deterministic:
            _, policy, _ = om.full_forward(stub, search, oo, color)
            kw["noise"] = dirichlet_sample(int(rng.integers(0, 1 << 30)), policy[:362]) if np.isfinite(policy[:362]).any() else np.zeros(362, np.float32)
            kw["choose_at"] = float(rng.random())
            kw["temperature"] = float(rng.choice([0.8, 0.05, 1.5]))
            kw["dirichlet_noise"] = float(rng.choice([0.25, 0.0, 0.7]))
        desc = f"case {k:3d} {where}, {len(plays)} stones, {kind} predictor, {kw['num_rollout']} rollouts, {kw['probes_per_round']} probes, " \
               f"search {search}, deterministic {deterministic}"
        try:
            tree, root = T.assert_same_search(stub, po, oo, color, **kw)
            evals_total += int(tree.total_count)
            line = desc + f": identical ({tree.total_count} visits, {int((tree.children()[0] > 0).sum())} children)"
        except AssertionError as exc:
            bad += 1
            line = desc + f": DIFFERENT {str(exc)[:200]}"
        print(line, flush=True)
        lines.append(line)
    lines.append(f"{args.cases} cases, {evals_total} visits compared, {bad} different ({time.time() - t0:.0f} s)")
    print(lines[-1])
    if args.out:
        with open(args.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
