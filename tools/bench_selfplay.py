#!/usr/bin/env python
"""Self-play throughput (BASELINE.json configs[2] / [3]): `--self-play N --num-rollout 800` with G concurrent games on
one B200 engine per process.  Fixed-duration sample of the full run (games with random-init weights go to the 722-ply
cap), stated in the output.

    python tools/bench_selfplay.py [--games 100] [--parallel 32] [--rollouts 800] [--probes 8] [--seconds 60]
    torchrun --nproc-per-node N tools/bench_selfplay.py ...      # one engine + one set of games per GPU, no collective

Prints ONE JSON line (rank 0): moves/s, NN evals/s, mean device batch, plus the host-only ceiling (same search with the
RandomPredictor, no device) so that the split between search cost and evaluation cost is visible.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def sample(net, *, games: int, parallel: int, rollouts: int, probes: int, seconds: float, threads: int, seed: int,
           ex_it: bool = False, device_features="queue", cache_capacity: int = 0, ex_it_rollouts: int = 0, device_priors=None,
           device_ladders=None, cache_shared: int = 0):
    """Runs a fixed-duration self-play sample on an existing engine; returns the driver's statistics.
    device_features: "queue" = the product path (leaf-batch queue, dg_selfplay_run_engine; device_priors None = chosen from
    the host threads per GPU), "prior" / True / False = the blocking predictor calls (priors on the device / planes on the
    device / planes on the host)."""
    from dream_go_b200 import mcts
    predictor = (mcts.EngineQueue(net, device_priors=device_priors, device_ladders=device_ladders) if device_features == "queue"
                 else mcts.EnginePriorPredictor(net) if device_features == "prior" else mcts.EngineRawPredictor(net) if device_features
                 else mcts.EnginePredictor(net))
    st, sgf = mcts.self_play(predictor, num_games=games, num_parallel=parallel, num_rollout=rollouts,
                             probes_per_round=probes, num_threads=threads, ex_it=ex_it, num_ex_it_rollout=ex_it_rollouts or rollouts, seed=seed,
                             max_seconds=seconds, cache_capacity=cache_capacity, cache_shared=cache_shared)
    return st, sgf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=100)
    ap.add_argument("--parallel", type=int, default=32)
    ap.add_argument("--rollouts", type=int, default=800)
    ap.add_argument("--probes", type=int, default=8)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--blocks", type=int, default=9)
    ap.add_argument("--weights", default=None, help="a dream_go.json weight file (loader.rs format) instead of seeded random-init weights")
    ap.add_argument("--ex-it", action="store_true")
    ap.add_argument("--ex-it-rollouts", type=int, default=0, help="`--num-ex-it-rollout` (default: same as --rollouts)")
    ap.add_argument("--no-host-sample", action="store_true", help="skip the host-only (RandomPredictor) sample")
    ap.add_argument("--host-only", action="store_true", help="RandomPredictor instead of the engine (no GPU needed)")
    ap.add_argument("--sgf-out", default=None, help="write the finished games' records (rank 0) to this file")
    ap.add_argument("--cache", type=int, default=0, help="entries of each game's transposition table (0 = none)")
    ap.add_argument("--spin-sync", action="store_true",
                    help="engine calls spin in the driver while they wait (default: they nap, DG_FLAG_BLOCKING_SYNC, and leave the cores to the search)")
    ap.add_argument("--blocking-sync", action="store_true", help="(the default now; kept for old command lines)")
    ap.add_argument("--device-priors", action="store_true", help="also build the priors on the device (dg_engine_forward_raw_prior)")
    ap.add_argument("--host-priors", action="store_true", help="build the priors on the host even with few host threads per GPU")
    ap.add_argument("--device-ladders", action="store_true", help="read the ladder planes on the device too (DG_RAW_DEVICE_LADDERS)")
    ap.add_argument("--host-ladders", action="store_true", help="read the ladder planes on the host even with very few host threads per GPU")
    ap.add_argument("--shared-cache", type=int, default=0, help="lock stripes of ONE transposition table of --cache entries shared by all games (0 = one table per game)")
    ap.add_argument("--blocking-calls", action="store_true",
                    help="drive the engine through the blocking predictor calls + one device thread per group (the round-1 driver) "
                         "instead of the leaf-batch queue")
    ap.add_argument("--no-graph", action="store_true", help="leaf batches are enqueued call by call (DG_FLAG_NO_GRAPH)")
    ap.add_argument("--host-features", action="store_true",
                    help="compute the feature planes on the host (compact positions) instead of on the device (raw positions)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from dream_go_b200 import mcts, nn, weights
    cores = os.cpu_count() or 1
    threads = args.threads or max(1, cores // world)

    kw = dict(num_games=args.games, num_parallel=args.parallel, num_rollout=args.rollouts, probes_per_round=args.probes,
              num_threads=threads, ex_it=args.ex_it, num_ex_it_rollout=args.rollouts, seed=20261017 + rank,
              max_seconds=args.seconds)
    if args.no_host_sample and not args.host_only:
        host = {"evals": 0, "moves": 0, "seconds": 1.0, "mean_batch": 0.0}
    else:
        host, _ = mcts.self_play(mcts.RandomPredictor(), **{**kw, "max_seconds": min(args.seconds, 15.0)})
    line = {"metric": "self_play_moves_per_s", "unit": "moves/s", "n_gpus": world, "higher_is_better": True,
            "config": {"workload": f"--self-play {args.games} --num-rollout {args.rollouts}, {args.parallel} concurrent games per GPU"
                                   + (" --ex-it" if args.ex_it else ""),
                       "sample": (f"fixed-duration sample: the first {args.seconds:.0f} s of the run (games with random-init weights reach the 722-stone cap)"
                                  if args.seconds > 0 else "all games played to the end"),
                       "probes_per_round": args.probes, "feature_planes": "host" if args.host_features else "device (csrc/features.cu)",
                       "priors": "device" if args.device_priors else "host" if args.host_priors or args.blocking_calls or args.host_features
                                 else "auto: host or device batch by batch, from the worker threads' load (DG_SELFPLAY_AUTO_PRIORS)",
                       "ladders": "device" if args.device_ladders or (not args.host_ladders and not args.blocking_calls and threads < 2) else "host",
                       "driver": "blocking predictor calls, one device thread per group" if args.blocking_calls or args.host_features
                                 else "leaf-batch queue (dg_selfplay_run_engine): one graph launch per batch, completion flag in pinned memory",
                       "host_threads_per_gpu": threads, "host_cores": cores,
                       "weights": f"{args.blocks} blocks x 128 filters, seeded random init", "data": "synthetic"},
            "host_only": {"evals_per_s": host["evals"] / host["seconds"], "moves_per_s": host["moves"] / host["seconds"],
                          "mean_batch": host["mean_batch"], "predictor": "RandomPredictor (no device)"}}
    if not args.host_only:
        from dream_go_b200 import shard
        shards = shard.Shards(backend="nccl")
        engine_kw = dict(device=local_rank, max_batch=512, num_workspaces=8,
                         flags=(0 if args.spin_sync else nn.FLAG_BLOCKING_SYNC) | (nn.FLAG_NO_GRAPH if args.no_graph else 0))
        if args.weights:
            net = nn.Network(**engine_kw)
            net.load_json(args.weights)
            line["config"]["weights"] = f"{args.weights} ({net.num_blocks()} blocks)"
        else:
            net = nn.Network.from_tensors(weights.synthetic_network(seed=20261017, num_blocks=args.blocks), **engine_kw)
        shards.barrier()
        t0 = time.perf_counter()
        st, sgf = sample(net, games=args.games, parallel=args.parallel, rollouts=args.rollouts, probes=args.probes,
                         seconds=args.seconds, threads=threads, seed=shards.seed(20261017), ex_it=args.ex_it,
                         device_features=("queue" if not (args.blocking_calls or args.host_features) else
                                          "prior" if args.device_priors else not args.host_features),
                         device_priors=True if args.device_priors else False if args.host_priors else None,
                         device_ladders=True if args.device_ladders else False if args.host_ladders else None,
                         cache_capacity=args.cache, cache_shared=args.shared_cache, ex_it_rollouts=args.ex_it_rollouts)
        wall = time.perf_counter() - t0
        tot = shards.selfplay_totals(st)
        shards.close()
        if args.sgf_out and rank == 0:
            with open(args.sgf_out, "w") as fh:
                fh.write("\n".join(sgf) + "\n")
        moves, evals, games, seconds, eval_s, rounds = (tot["moves"], tot["evals"], tot["games_finished"], tot["seconds"],
                                                        tot["predictor_seconds"], tot["rounds"])
        line.update({"value": moves / seconds, "nn_evals_per_s": evals / seconds, "games_finished": games, "moves": moves,
                     "evals": evals, "seconds": seconds, "mean_batch": evals / max(rounds, 1), "cache_hits_rank0": st.get("cache_hits", 0), "searches_rank0": st.get("searches", 0),
                     "ex_it_expansions_rank0": (st.get("searches", 0) - st.get("moves", 0)) if args.ex_it else 0, "cache_capacity_per_game": args.cache,
                     "device_busy_frac": eval_s / (seconds * world), "wall_s": wall,
                     "first_game": sgf[0][:200] if sgf else None})
        net.close()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
