#!/usr/bin/env python
"""BASELINE.json configs[0]: libdg_go 19x19 feature-plane extract + legal-move generation, batch = 32, CPU only.

Positions: (i) every ply of the 99 fixture games (real), (ii) seeded uniform-random legal
playouts.  For each board: 361 x `Board::is_valid(to_move)` + V1 features (identity symmetry).  Times the PRODUCT's host
code (csrc/go_board.h: compact positions + legal mask from one pass) against the ORACLE restatement of the reference
(oracle/dg_oracle_go.cpp: `get_features::<HWC, f16>` + 361 x `is_valid`), single thread and all host cores.

    python tools/bench_go.py [--positions 4096]        -> one JSON line
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def collect_positions(limit: int, seed: int = 20261017):
    """Returns lists of (product board, oracle board, to_move) for `limit` corpus plies + `limit` random-playout plies."""
    from dream_go_b200 import go as pgo
    from oracle import go as ogo
    ogo.use_default_zobrist()
    games = ogo.load_games()
    rng = np.random.default_rng(seed)
    out = {"corpus": [], "random": []}
    picks = sorted(rng.choice(sum(len(m) for _, m, _ in games), size=limit, replace=False).tolist())
    at, pi = 0, 0
    for colors, moves, komi in games:
        po, oo = pgo.Board(komi), ogo.Board(komi)
        for c, m in zip(colors, moves):
            if pi < len(picks) and picks[pi] == at:
                out["corpus"].append((po.clone(), oo.clone(), int(c)))
                pi += 1
            at += 1
            if m < 361:
                po.place_index(int(c), int(m))
                oo.place_index(int(c), int(m))
    while len(out["random"]) < limit:
        po, oo = pgo.Board(7.5), ogo.Board(7.5)
        color = 1
        for ply in range(int(rng.integers(0, 300))):
            legal = np.flatnonzero(po.legal_moves(color))
            if len(legal) == 0:
                break
            m = int(rng.choice(legal))
            po.place_index(color, m)
            oo.place_index(color, m)
            color = 3 - color
        out["random"].append((po, oo, color))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--positions", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    from dream_go_b200 import go as pgo
    from oracle import go as ogo
    sets = collect_positions(args.positions)
    cores = os.cpu_count() or 1
    line = {"metric": "feature_extract_positions_per_s", "unit": "positions/s", "higher_is_better": True,
            "config": {"workload": "libdg_go 19x19 feature-plane extract + legal-move gen, batch=32 (BASELINE.json configs[0])",
                       "positions_per_set": args.positions, "host_cores": cores}}
    for name, items in sets.items():
        boards = [p for p, _, _ in items]
        tm = np.array([c for _, _, c in items], np.uint8)
        res = {}
        for label, threads in (("1_thread", 1), (f"{cores}_threads", cores)):
            batches = [pgo.PreparedBatch(boards[i:i + args.batch], tm[i:i + args.batch], legal=True, threads=threads)
                       for i in range(0, len(boards), args.batch)]
            for b in batches:
                b.run()
            t0 = time.perf_counter()
            for b in batches:
                b.run()
            res[label] = len(boards) / (time.perf_counter() - t0)
        # oracle: single thread, one board at a time (the reference extracts per leaf on the probing thread)
        sample = items[:min(len(items), 512)]
        nodes0 = ogo.lib().dgo_ladder_nodes()
        t0 = time.perf_counter()
        for _, oo, c in sample:
            oo.features(c)
            oo.legal_mask(c)
        dt = time.perf_counter() - t0
        res["oracle_1_thread"] = len(sample) / dt
        res["oracle_ladder_nodes_per_position"] = (ogo.lib().dgo_ladder_nodes() - nodes0) / len(sample)
        res["speedup_1_thread"] = res["1_thread"] / res["oracle_1_thread"]
        line[name] = res
    line["value"] = line["corpus"][f"{cores}_threads"]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
