#!/usr/bin/env python
"""Soak run of the host Go code against the oracle restatement of libdg_go at a scale the test suite cannot afford (the same
checks as tests/test_go_parity.py): hashes, legal moves (with super-ko) and all 32 feature planes bit for bit on long random
playouts (captures, kos, refilled areas), the raw positions' ladder planes on every ply, unconditional life / scoring
candidates / territory on settled positions.  CPU only.

    python tools/soak_go_parity.py [--playouts 300] [--out profiles/r02_soak_go_vs_oracle.log]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import test_go_parity as T                                  # noqa: E402
from dream_go_b200 import go as pgo                         # noqa: E402
from oracle import go as ogo                                # noqa: E402

BLACK, WHITE = 1, 2


def bits(words):
    return np.unpackbits(np.asarray(words, "<u4").view(np.uint8), bitorder="little")[:361]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--playouts", type=int, default=300)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    ogo.use_default_zobrist()
    lines = []

    def say(s):
        print(s, flush=True)
        lines.append(s)

    # 1. replays: hashes, legal masks, 32 planes at every ply
    t0 = time.time()
    positions = 0
    for seed in range(1000, 1000 + args.playouts):
        plies = 200 + 37 * (seed % 30)                       # 200 .. 1,273 plies
        komi = [7.5, 6.5, 0.5, -3.5, 11.5][seed % 5]
        colors, moves = T.random_playout(seed, plies, komi=komi, pass_rate=[0.0, 0.02, 0.1][seed % 3])
        T.assert_same_replay(colors, moves, komi)
        positions += plies
    say(f"replays: {args.playouts} random playouts, {positions} positions: hashes, legal moves and all 32 feature planes identical ({time.time() - t0:.0f} s)")

    # 2. raw positions: ladder planes (host reader) on every ply, both colours to move
    t0 = time.time()
    checked = captures = escapes = 0
    for seed in range(2000, 2000 + max(1, args.playouts // 6)):
        colors, moves = T.random_playout(seed, 330, pass_rate=0.0)
        po, oo = pgo.Board(7.5), ogo.Board(7.5)
        for c, m in zip(colors, moves):
            if m < 361:
                po.place_index(int(c), int(m))
                oo.place_index(int(c), int(m))
            for to_move in (BLACK, WHITE):
                raw = po.raw_position(to_move)[0]
                want = oo.features(to_move).reshape(361, 32).astype(np.float32)
                cap, esc = bits(raw["ladder_capture"]), bits(raw["ladder_escape"])
                assert (cap == (want[:, 30] != 0)).all() and (esc == (want[:, 31] != 0)).all(), (seed, to_move)
                captures += int(cap.sum())
                escapes += int(esc.sum())
                checked += 1
    say(f"raw positions: {checked} positions (both colours to move): {captures} ladder captures, {escapes} ladder escapes identical ({time.time() - t0:.0f} s)")

    # 3. unconditional life, scoring candidates, scorable, territory
    t0 = time.time()
    n = scorable = 0
    for seed in range(100, 100 + 2 * args.playouts):
        po, oo = pgo.Board(6.5), ogo.Board(6.5)
        for c, x, y in T.settled_position(seed):
            if oo.at(x, y) == 0 and oo.is_valid(c, x, y):
                po.place(c, x, y)
                oo.place(c, x, y)
        assert (po.stones() == oo.stones()).all()
        for color in (BLACK, WHITE):
            assert (po.benson(color) == oo.benson(color)).all(), seed
            for kind in (0, 1):
                assert (po.policy_candidates(color, kind) == oo.policy_candidates(color, kind)).all(), seed
        assert po.is_scorable() == oo.is_scorable()
        assert (po.territory() == oo.territory()).all(), seed
        scorable += int(oo.is_scorable())
        n += 1
    # late random positions as well (nothing settled: the early exits)
    for seed in range(3000, 3000 + max(1, args.playouts // 6)):
        colors, moves = T.random_playout(seed, 420, pass_rate=0.0)
        po, oo = pgo.Board(7.5), ogo.Board(7.5)
        for k, (c, m) in enumerate(zip(colors, moves)):
            if m < 361:
                po.place_index(int(c), int(m))
                oo.place_index(int(c), int(m))
            if k % 7 == 0:
                for color in (BLACK, WHITE):
                    assert (po.benson(color) == oo.benson(color)).all(), (seed, k)
                    assert (po.policy_candidates(color, 1) == oo.policy_candidates(color, 1)).all(), (seed, k)
                assert (po.territory() == oo.territory()).all(), (seed, k)
                n += 1
    # fully settled boards: black fills the lines below a random border line, white the lines above, each side with a random number of
    # one-point eyes (two or more: alive, everything is settled and the board is scorable; fewer: it is not)
    rng = np.random.default_rng(9)
    for k in range(max(4, args.playouts // 5)):
        border = int(rng.integers(4, 15))
        holes = {}
        for color, rows in ((BLACK, range(0, border)), (WHITE, range(border, 19))):
            want, pts = int(rng.integers(0, 5)), []
            for _ in range(200):
                if len(pts) == want:
                    break
                x, y = int(rng.integers(0, 19)), int(rng.choice(list(rows)))
                if (color == BLACK and y == border - 1) or (color == WHITE and y == border):
                    continue                                     # an eye on the border line would touch the other colour
                if all(abs(x - a) + abs(y - b) > 1 for a, b in pts):
                    pts.append((x, y))
            holes[color] = pts
        po, oo = pgo.Board(7.5), ogo.Board(7.5)
        for y in range(19):
            for x in range(19):
                color = BLACK if y < border else WHITE
                if (x, y) not in holes[color] and oo.is_valid(color, x, y):
                    po.place(color, x, y)
                    oo.place(color, x, y)
        assert (po.stones() == oo.stones()).all()
        for color in (BLACK, WHITE):
            assert (po.benson(color) == oo.benson(color)).all(), k
            for kind in (0, 1):
                assert (po.policy_candidates(color, kind) == oo.policy_candidates(color, kind)).all(), k
        assert po.is_scorable() == oo.is_scorable(), k
        assert (po.territory() == oo.territory()).all(), k
        scorable += int(oo.is_scorable())
        n += 1
    say(f"unconditional life: {n} positions ({scorable} scorable): Benson sets, candidates of both search kinds, is_scorable, territory identical "
        f"({time.time() - t0:.0f} s)")
    # 4. priors (create_initial_policy + add_valid_candidates + normalize_policy) through dg_board_prior, bit for bit: random and
    #    symmetric positions (candidates folded onto orbit representatives), every symmetry of the network's policy, policies with
    #    zeros, tiny entries and all zeros
    t0 = time.time()
    rng = np.random.default_rng(4)
    n = folded = 0
    for k in range(max(2, args.playouts // 10)):
        po, oo = pgo.Board(7.5), ogo.Board(7.5)
        if k % 3 == 0:
            group = [[0, 1], [0, 2], [0, 3], [0, 4], [0, 6], [0, 1, 2, 6], [0, 3, 4, 6], list(range(8))][int(rng.integers(0, 8))]
            taken = set()
            for _ in range(int(rng.integers(0, 16))):
                p0, c = int(rng.integers(0, 361)), int(rng.integers(1, 3))
                orbit = sorted({ogo.symmetry_apply(t, p0) for t in group})
                if not taken & set(orbit):
                    taken |= set(orbit)
                    for q in orbit:
                        po.place_index(c, q)
                        oo.place_index(c, q)
        else:
            colors, moves = T.random_playout(int(rng.integers(0, 1 << 20)), int(rng.integers(0, 300)), pass_rate=0.0)
            for c, m in zip(colors, moves):
                if m < 361:
                    po.place_index(int(c), int(m))
                    oo.place_index(int(c), int(m))
        assert (po.stones() == oo.stones()).all()
        folded += any(oo.is_symmetric(t) for t in range(1, 8))
        for to_move in (BLACK, WHITE):
            for symmetry in range(8):
                logits = rng.normal(size=362).astype(np.float32) * float(rng.choice([0.5, 3.0, 12.0]))
                policy = (np.exp(logits - logits.max()) / np.exp(logits - logits.max()).sum()).astype(np.float16)
                kind = int(rng.integers(0, 4))
                if kind == 1:
                    policy[rng.random(362) < 0.5] = 0
                elif kind == 2:
                    policy[:] = 0
                sum_to = float(rng.choice([1.0, 0.125]))
                want = T.oracle_prior(oo, to_move, policy, symmetry, sum_to)
                got = po.prior(to_move, policy, symmetry, sum_to)
                f = np.isfinite(want)
                assert (np.isfinite(got) == f).all() and (got[f].view(np.uint32) == want[f].view(np.uint32)).all(), (k, to_move, symmetry)
                n += 1
    say(f"priors: {n} priors on {max(2, args.playouts // 10)} positions ({folded} with a symmetry): candidates and values bit-identical ({time.time() - t0:.0f} s)")
    # 5. exact liberty counts: `get_n_liberty` of every stone, `get_n_liberty_if` of both colours at every point where the move
    #    is legal without the ko rule (the planes only pin the counts up to 6)
    t0 = time.time()
    n = q = 0
    for seed in range(5000, 5000 + max(2, args.playouts // 20)):
        colors, moves = T.random_playout(seed, 400, pass_rate=0.0)
        po, oo = pgo.Board(7.5), ogo.Board(7.5)
        for k, (c, m) in enumerate(zip(colors, moves)):
            if m < 361:
                po.place_index(int(c), int(m))
                oo.place_index(int(c), int(m))
            if k % 5:
                continue
            for y in range(19):
                for x in range(19):
                    if oo.at(x, y):
                        assert po.get_n_liberty(x, y) == oo.get_n_liberty(x, y), (seed, k, x, y)
                        q += 1
                    else:
                        for col in (BLACK, WHITE):
                            if oo.is_valid_fast(col, x, y):
                                assert po.get_n_liberty_if(col, x, y) == oo.get_n_liberty_if(col, x, y), (seed, k, x, y, col)
                                q += 1
            n += 1
    say(f"liberties: {q} exact liberty counts (stones, and both colours' liberties-if-played at every playable point) on {n} positions identical "
        f"({time.time() - t0:.0f} s)")
    say("ALL IDENTICAL")
    if args.out:
        with open(args.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
