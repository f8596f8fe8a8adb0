#!/usr/bin/env python
"""`dream_go --self-play N [--ex-it]` on this engine: plays N games and prints one SGF record per finished game to stdout
(`;B[xy]TR[..]TV[n]P[b85 fp16 visit distribution]V[..]`, what contrib/trainer reads), a `.` per game to stderr as the
reference's main.rs:72-83 does.  Flags follow src/libdg_utils/config.rs: --num-rollout, --num-ex-it-rollout, --num-games
(concurrent games), --num-threads, --batch-size is not needed (the leaf-batch queue sizes its batches from the games).

    python tools/self_play.py --self-play 100 [--ex-it] [--num-rollout 800] [--num-games 32] [--weights dream_go.json]
    python tools/self_play.py --self-play 4 --host-only           # the RandomPredictor instead of a device (no GPU needed)

Every GPU of the process gets an engine (`Device::all()`, predictors/nn.rs:84-92); the games are dealt to them round-robin and
the host threads are shared (dg_selfplay_run_engine).  This is a thin front end over the C ABI -- the reference's CLI, GTP and
time control are out of scope.
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--self-play", type=int, required=True, metavar="N", help="number of games to generate")
    ap.add_argument("--ex-it", action="store_true", help="expert iteration: 5 %% of the eligible moves get a second, deeper search")
    ap.add_argument("--num-rollout", type=int, default=800)
    ap.add_argument("--num-ex-it-rollout", type=int, default=800)
    ap.add_argument("--num-games", type=int, default=32, help="games played concurrently (per process)")
    ap.add_argument("--num-threads", type=int, default=0, help="host threads (0 = all cores)")
    ap.add_argument("--probes", type=int, default=8, help="leaves per tree and device batch")
    ap.add_argument("--weights", default=None, help="dream_go.json (default: the reference's search order, network.rs:92-124)")
    ap.add_argument("--blocks", type=int, default=9, help="with --random-weights: residual blocks of the seeded random-init network")
    ap.add_argument("--random-weights", action="store_true", help="seeded random-init weights when no weight file is at hand")
    ap.add_argument("--cache", type=int, default=200000, help="entries of the process-wide transposition table (predictors/nn.rs:48-50); 0 = none")
    ap.add_argument("--seed", type=int, default=None, help="default: from the clock, as the reference's thread_rng")
    ap.add_argument("--host-only", action="store_true", help="RandomPredictor instead of the engine")
    args = ap.parse_args()

    from dream_go_b200 import mcts, nn, weights
    seed = args.seed if args.seed is not None else (int.from_bytes(os.urandom(4), "little") | 1)
    kw = dict(num_games=args.self_play, num_parallel=args.num_games, num_rollout=args.num_rollout, probes_per_round=args.probes,
              num_threads=args.num_threads, ex_it=args.ex_it, num_ex_it_rollout=args.num_ex_it_rollout, seed=seed,
              cache_capacity=args.cache, cache_shared=64 if args.cache > 0 else 0,
              sgf_capacity=max(1 << 24, args.self_play * (1 << 21)))
    nets = []
    if args.host_only:
        predictor = mcts.RandomPredictor()
    else:
        devices = max(1, nn.lib().dg_device_count())
        per_engine = (args.num_games + devices - 1) // devices
        engine_kw = dict(max_batch=max(256, min(512, per_engine * max(8, args.probes))), num_workspaces=4, flags=nn.FLAG_BLOCKING_SYNC)
        for device in range(devices):
            if args.weights:
                net = nn.Network(device=device, **engine_kw)
                net.load_json(args.weights)
            elif args.random_weights:
                net = nn.Network.from_tensors(weights.synthetic_network(seed=20261017, num_blocks=args.blocks), device=device, **engine_kw)
            else:
                net = nn.Network.new(device=device, **engine_kw)
                if net is None:
                    print("no weights file found (dream_go.json); give --weights or --random-weights", file=sys.stderr)
                    return 1
            nets.append(net)
        predictor = mcts.EngineQueue(nets)
    stats, games = mcts.self_play(predictor, **kw)
    for sgf in games:
        sys.stdout.write(sgf + "\n")
        sys.stderr.write(".")
    sys.stderr.write(f"\n{int(stats['games_finished'])} games, {int(stats['moves'])} moves, {int(stats['evals'])} evaluations "
                     f"({int(stats['cache_hits'])} answered by the table) in {stats['seconds']:.1f} s\n")
    for net in nets:
        net.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
