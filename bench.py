#!/usr/bin/env python
"""Benchmark of the self-play NN evaluation path (BASELINE.json configs[1]):
9-block x 128-filter residual tower + heads, batch 256, synthetic weights and positions.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one forward of one batch (256 positions) through the whole network.  Prints ONE
JSON line (rank 0).  See DESIGN.md "Measurement" for every field.

* value      : NN evals/s with the batch already resident in HBM; every step timed with its own
               CUDA-event pair on the engine's stream, L2 flushed (256 MiB memset) between steps.
* e2e        : the same through the reference-facing C-ABI call `dg_engine_forward_f16`
               (== dg_nn::forward) from pinned HOST buffers, H2D + D2H inside the timed region.
* roofline   : dominant kernel = tower_kernel (ONE persistent launch per step that runs the up-sampling
               layer and all 18 residual 3x3 128->128 convolutions); tensor bound.  `achieved` is timed
               live on the 18 residual convolutions alone (a second pass of tower-only launches).
* sustained  : the same resident step repeated for >= 2 s of device time (the power-capped regime), against
               MEASURED_PEAKS.json:bf16_tflops_sustained.
* self_play  : the metric's other half -- moves/s and evals/s of `--self-play --num-rollout 800` for configs[2] (32
               concurrent games), configs[3]'s shape (64 per GPU) and 128 per GPU, and at N = 1 the same loop on the
               reference's cuDNN evaluation path under the reference's batching rules (`cudnn_reference`).
* cudnn_baseline / vs_cudnn : the reference's 25-call cuDNN forward restated on this box's cuDNN -- the GPU-vs-GPU ratio.
* cpu_baseline / --impl reference : the CPU oracle port of dg_nn::forward on the host cores (a GPU-vs-CPU figure).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 256
NUM_BLOCKS = 9
METRIC = "nn_evals_per_s"
UNIT = "evals/s"
FLOP_PER_EVAL = 2 * 976_681_890
TOWER_CONV_FLOP_PER_POS = 2 * 361 * 1152 * 128    # one 3x3 128->128 convolution, algorithmic
WORKLOAD = "residual tower 9-block x 128-filter forward, batch=256 (BASELINE.json configs[1])"


def workload_config(world: int):
    """The workload both arms run (BASELINE.json configs[1]); the same dictionary in both arms' lines."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "blocks": NUM_BLOCKS, "filters": 128,
            "inputs": "iid Bernoulli(0.2) fp16 features (dg_tests/benches/batch_sizes.rs:44-49), seeded He-init weights",
            "parallelism": f"{world} independent engine shard(s), no collective",
            "l2": "device arm: flushed between steps (256 MiB memset outside the per-step event pairs); CPU arm: not applicable"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured"
    return 1590.0, 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)
        return False

    def summary(self):
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, flag in zip(names, r[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        top = sorted(sm)[len(sm) // 2:]             # upper half = samples under load
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_rate(seconds_target: float, steps: int = 1, warmup: int = 0):
    """Times the CPU oracle port of dg_nn::forward on a bounded sample; returns (evals/s, info)."""
    from dream_go_b200 import weights
    from oracle import oracle
    net = weights.synthetic_network(seed=20261017, num_blocks=NUM_BLOCKS)
    onet = oracle.OracleNetwork(net)
    probe = weights.bernoulli_features(2, seed=11)
    t0 = time.perf_counter()
    onet.forward(probe)
    per_pos = (time.perf_counter() - t0) / 2
    sample = int(max(2, min(BATCH, seconds_target / max(per_pos, 1e-6) / max(steps + warmup, 1))))
    feats = weights.bernoulli_features(sample, seed=12)
    for _ in range(warmup):
        onet.forward(feats)
    t0 = time.perf_counter()
    for _ in range(steps):
        onet.forward(feats)
    dt = time.perf_counter() - t0
    return sample * steps / dt, {"cores": oracle.num_threads(), "kind": "port", "seconds": dt,
                                 "sample": f"{sample} of the {BATCH} positions of one batch x {steps} step(s), same weights/inputs distribution"}


def host_feature_rates(positions: int = 512):
    """BASELINE.json configs[0] (CPU only): V1 feature planes + legal moves for batches of 32 boards -- the product's
    host code next to the oracle port of the reference's libdg_go, on real game positions (tools/bench_go.py)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_go
        from dream_go_b200 import go as pgo
        from oracle import go as ogo
        items = bench_go.collect_positions(positions)["corpus"]
        boards = [p for p, _, _ in items]
        tm = np.array([c for _, _, c in items], np.uint8)
        cores = os.cpu_count() or 1
        out = {"unit": "positions/s", "workload": "feature-plane extract + legal-move gen, batch=32, fixture-game positions", "cores": cores}
        for label, threads in (("product_1_thread", 1), ("product_all_cores", cores)):
            batches = [pgo.PreparedBatch(boards[i:i + 32], tm[i:i + 32], legal=True, threads=threads) for i in range(0, len(boards), 32)]
            for b in batches:
                b.run()
            t0 = time.perf_counter()
            for rep in range(4):
                for b in batches:
                    b.run()
            out[label] = 4 * len(boards) / (time.perf_counter() - t0)
        t0 = time.perf_counter()
        for _, oo, c in items:
            oo.features(c)
            oo.legal_mask(c)
        out["oracle_port_1_thread"] = len(items) / (time.perf_counter() - t0)
        return out
    except Exception as exc:   # noqa: BLE001
        return {"unavailable": repr(exc)[:200]}


def cudnn_baseline(tensors, feats, steps: int):
    """The reference's own GPU path restated call for call on the image's cuDNN (baseline/cudnn_ref.cu),
    same weights, same positions, same box: device-resident and end-to-end (blocking H2D/D2H) rates."""
    try:
        from baseline import cudnn_ref
        ref = cudnn_ref.CudnnNetwork(tensors, BATCH)
        host = np.array(feats)                      # the reference passes a plain (pageable) slice
        ref.forward(host)
        ref.time_resident(5)
        ms = ref.time_resident(steps) / steps
        t0 = time.perf_counter()
        for _ in range(steps):
            ref.forward(host)
        e2e = (time.perf_counter() - t0) / steps
        info = ref.info
        ref.close()
        return {"value": BATCH / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "e2e": BATCH / e2e, "kind": "cuDNN 9.10.2 legacy API, 25 launches/forward (graph.rs:123-158)",
                "algos": info[:160]}
    except Exception as exc:   # noqa: BLE001
        return {"unavailable": repr(exc)[:200]}


def run_reference(args, rank: int):
    """`--impl reference`: the reference's CPU-side restatement (oracle port; the Rust + cuDNN
    reference cannot be built in this image) on the host cores."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to every rank; this arm is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    rate, info = oracle_rate(seconds_target=60.0, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
    per_step_ms = 1e3 * info["seconds"] / max(args.steps, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": workload_config(args.gpus),       # the same dictionary as the device arm's line (the CPU arm's sample is below)
        "sample": info["sample"],
        "note": "CPU restatement of dg_nn::forward on the host cores of rank 0 (the Rust + cuDNN reference cannot be built here)",
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def self_play_summary(self_play: dict) -> dict:
    """The metric's other half in a form that survives the driver's parsing (sub-keys of `e2e` are kept; unknown top-level keys
    are not) and its 1,500-character tail of stdout (the last key of the line): [moves/s, NN evals/s] per sample."""
    pair = lambda d: [round(d["moves_per_s"], 1), round(d["nn_evals_per_s"])]
    out = {"unit": "[moves/s, NN evals/s], --self-play --num-rollout 800, whole job", "host_threads_per_gpu": self_play["host_threads_per_gpu"],
           "configs2_32_games_per_gpu": pair(self_play["configs2"]), "configs3_shape_64_games_per_gpu": pair(self_play["configs3_shape"]),
           "games128_per_gpu": pair(self_play["games128"])}
    if "games128_shared_table" in self_play:
        out["games128_shared_table"] = pair(self_play["games128_shared_table"])
    ref = self_play.get("cudnn_reference") or {}
    for bs in (16, 32):
        if f"batch{bs}" in ref:
            out[f"cudnn_reference_32_games_batch{bs}"] = pair(ref[f"batch{bs}"])
    if "batch16" in ref and ref["batch16"]["moves_per_s"] > 0:
        out["configs2_vs_cudnn_reference_batch16"] = round(self_play["configs2"]["moves_per_s"] / ref["batch16"]["moves_per_s"], 2)
    return out


def self_play_samples(args, tensors, shards, rank: int, world: int, local_rank: int):
    """Fixed-duration samples of `--self-play --num-rollout 800` from the empty board (random-init weights: games run to
    the 722-ply cap) through the product path -- leaf-batch queue, planes / legal moves (/ priors) on the device:
    configs[2] (32 concurrent games on one GPU), configs[3]'s shape (64 concurrent games per GPU) and 128 per GPU; at
    N = 1 also the SAME self-play loop on the reference's evaluation path -- cuDNN restatement of dg_nn::forward under the
    reference's batching rules (<= --batch-size leaves per batch, <= 2 batches in flight per GPU, one leaf per probe,
    fp16 feature tensors from the host; pool/batch.rs:98-123, predictors/nn.rs:64-67) -- and the host half alone."""
    from dream_go_b200 import mcts, nn
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_selfplay
    threads = max(1, (os.cpu_count() or 1) // world)
    secs = args.self_play_seconds
    # its own engine, one workspace per group of games; waits nap instead of spinning on a core
    sp_net = nn.Network.from_tensors(tensors, device=local_rank, max_batch=2 * BATCH, num_workspaces=4, flags=nn.FLAG_BLOCKING_SYNC)
    out = {"unit": "moves/s, evals/s", "rollouts": 800, "probes_per_round": 8, "host_threads_per_gpu": threads, "host_cores": os.cpu_count(),
           "driver": "dg_selfplay_run_engine (leaf-batch queue, graph launch per batch, device planes)",
           "priors": "host or device, decided batch by batch from the worker threads' load (DG_SELFPLAY_AUTO_PRIORS)",
           "ladders": "device" if threads < 2 else "host", "sample_seconds": secs}
    for key, games in (("configs2", 32), ("configs3_shape", 64), ("games128", 128)):
        shards.barrier()
        st, _ = bench_selfplay.sample(sp_net, games=100000, parallel=games, rollouts=800, probes=8, seconds=secs, threads=threads,
                                      seed=20261017 + rank)
        tot = shards.selfplay_totals(st)
        out[key] = {"concurrent_games_per_gpu": games, "moves_per_s": tot["moves_per_s"], "nn_evals_per_s": tot["nn_evals_per_s"],
                    "mean_device_batch": tot["mean_device_batch"], "device_busy_frac": tot["predictor_seconds"] / (tot["seconds"] * world)}
    # the reference keeps ONE 200,000-entry transposition table for the whole process (predictors/nn.rs:29-82): the same here
    # (games then depend on each other's timing, as the reference's do); evaluations answered by the table are not counted
    shards.barrier()
    st, _ = bench_selfplay.sample(sp_net, games=100000, parallel=128, rollouts=800, probes=8, seconds=secs, threads=threads,
                                  seed=20261017 + rank, cache_capacity=200000, cache_shared=64)
    tot = shards.selfplay_totals(st)
    out["games128_shared_table"] = {"concurrent_games_per_gpu": 128, "moves_per_s": tot["moves_per_s"], "nn_evals_per_s": tot["nn_evals_per_s"],
                                    "table": "one table of 200,000 entries in 64 lock stripes per process (cache_shared)",
                                    "table_hits_rank0": st.get("cache_hits", 0), "evals_rank0": st.get("evals", 0)}
    sp_net.close()
    if rank == 0 and world == 1:
        # the reference's evaluation path behind the same loop
        ref = {}
        try:
            from baseline import cudnn_ref
            for bs in (16, 32):
                pred = cudnn_ref.ReferencePredictor(tensors, batch_size=bs, lanes=2, device=local_rank)
                # 2 groups of 16 games, bs / 16 leaves per game and round: every group round is one full batch, two in flight
                st, _ = mcts.self_play(pred, num_games=100000, num_parallel=32, num_rollout=800, probes_per_round=max(1, bs // 16),
                                       num_threads=threads, seed=20261017, max_seconds=secs, num_groups=2)
                ps = pred.stats()
                pred.close()
                ref[f"batch{bs}"] = {"moves_per_s": st["moves"] / st["seconds"], "nn_evals_per_s": st["evals"] / st["seconds"],
                                     "mean_cudnn_batch": ps["mean_batch"], "concurrent_games": 32, "moves": st["moves"], "evals": st["evals"],
                                     "seconds": st["seconds"]}
            ref["rules"] = "cuDNN 9.10.2 forward, <= batch-size leaves per batch, <= 2 batches in flight, 1 leaf per probe, fp16 features from pageable host memory"
        except Exception as exc:   # noqa: BLE001
            ref = {"unavailable": repr(exc)[:200]}
        out["cudnn_reference"] = ref
        if "batch16" in ref and ref["batch16"]["nn_evals_per_s"] > 0 and ref["batch32"]["nn_evals_per_s"] > 0:
            # same loop, same rollouts per move: the evaluation rates are the stable ratio, the move rates the headline
            out["configs2_vs_cudnn_reference"] = {
                f"batch{bs}": {"evals": out["configs2"]["nn_evals_per_s"] / ref[f"batch{bs}"]["nn_evals_per_s"],
                               "moves": out["configs2"]["moves_per_s"] / ref[f"batch{bs}"]["moves_per_s"] if ref[f"batch{bs}"]["moves_per_s"] > 0 else None}
                for bs in (16, 32)}
        # the host half alone: same search and feature code, RandomPredictor instead of the device
        ho, _ = mcts.self_play(mcts.RandomPredictor(), num_games=100000, num_parallel=128, num_rollout=800,
                               probes_per_round=8, num_threads=threads, seed=20261017, max_seconds=4.0)
        out["host_only"] = {"nn_evals_per_s": ho["evals"] / ho["seconds"], "moves_per_s": ho["moves"] / ho["seconds"], "threads": threads}
    return out


def run_ours(args, rank: int, world: int, local_rank: int):
    from dream_go_b200 import nn, shard, weights

    shards = shard.Shards(backend="nccl")       # plumbing only: start barrier + adding up the shards' counters
    barrier, max_over_ranks, sum_over_ranks = shards.barrier, shards.max, shards.sum

    tensors = weights.synthetic_network(seed=20261017, num_blocks=NUM_BLOCKS)
    net = nn.Network.from_tensors(tensors, device=local_rank, max_batch=BATCH, num_workspaces=3)
    feats = net.pinned((BATCH, 361, 32), np.float16)
    feats[...] = weights.bernoulli_features(BATCH, seed=1000 + rank)       # each rank (shard) has its own positions
    value = net.pinned((BATCH,), np.float16)
    policy = net.pinned((BATCH, 362), np.float16)

    # ---- warm-up (also makes the batch resident) -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        net.forward_into(feats, value, policy)
    assert np.isfinite(policy.astype(np.float32)).all() and abs(float(policy[0].astype(np.float32).sum()) - 1.0) < 2e-2

    # ---- device-resident timing -------------------------------------------------------------------
    barrier()
    net.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    ms_total, tower_ms, launches = net.time_resident(BATCH, args.steps, tower=True, flush_l2=True)
    net.synchronize()
    barrier()
    ms_total = max_over_ranks(ms_total)
    ms_per_step = ms_total / args.steps
    value_evals = world * BATCH * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI from pinned host buffers --------------------------------------
    # (a) one caller, strictly serial H2D -> kernels -> D2H per call
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net.forward_into(feats, value, policy)
    e2e_single_s = max_over_ranks(time.perf_counter() - t0)
    # (b) two callers, as the reference drives one device (`max_num_threads() = 2 x device_count`,
    #     predictors/nn.rs:64-67): each call is still a blocking H2D + forward + D2H of its own batch from its own
    #     pinned buffers, but one caller's copies overlap the other's kernels.  Native host threads inside the library
    #     (dg_engine_time_e2e) so that the Python GIL is not part of the measurement.  Same total number of steps.
    barrier()
    e2e_s = max_over_ranks(net.time_e2e(feats, args.steps, callers=2))
    e2e_evals = world * BATCH * args.steps / e2e_s
    barrier()
    e2e3_s = max_over_ranks(net.time_e2e(feats, args.steps, callers=3))
    packed = net.pinned((BATCH,), nn.PACKED_DTYPE)
    packed[...] = nn.pack_positions(feats)
    net.forward_into(packed, value, policy, packed=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        net.forward_into(packed, value, policy, packed=True)
    e2e_packed_s = max_over_ranks(time.perf_counter() - t0)
    clocks.__exit__()

    # ---- sustained: the same resident step for >= 2 s of device time (power-capped clocks), L2 flushed between steps -------
    sustained = None
    if args.sustained_seconds > 0:
        steps_sus = max(args.steps, int(args.sustained_seconds / (ms_per_step * 1e-3)) + 1)
        barrier()
        ms_sus, tower_ms_sus, _ = net.time_resident(BATCH, steps_sus, tower=True, flush_l2=True)
        ms_sus = max_over_ranks(ms_sus)
        sustained = {"steps": steps_sus, "timed_seconds": ms_sus * 1e-3, "value": world * BATCH * steps_sus / (ms_sus * 1e-3), "unit": UNIT,
                     "tower_us_per_launch": tower_ms_sus * 1e3 / steps_sus}

    # ---- the metric's other half: self-play moves/s through the whole hot path (BASELINE.json configs[2]/[3]) -----
    self_play = None
    if args.self_play_seconds > 0:
        self_play = self_play_samples(args, tensors, shards, rank, world, local_rank)

    shards.close()
    if rank != 0:
        return
    peak_tf, _peak_gbs, peak_kind = measured_peaks()
    tower_launch_s = tower_ms * 1e-3 / args.steps                      # one tower-only launch = 18 convolutions
    achieved_tf = 2 * NUM_BLOCKS * BATCH * TOWER_CONV_FLOP_PER_POS / tower_launch_s / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get("tower_kernel_dram_bytes_per_launch")
    peak_sus = None
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        with open(ppath) as fh:
            peak_sus = json.load(fh).get("bf16_tflops_sustained")
    line = {
        "metric": METRIC, "value": value_evals, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
    }
    if self_play is not None:          # the metric's other half, next to the headline
        line["self_play"] = self_play
    line.update({
        "config": workload_config(world),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_evals, "unit": UNIT, "h2d_bytes_per_step": BATCH * 11552 * 2, "d2h_bytes_per_step": BATCH * 363 * 2,
                "call": "dg_engine_forward_f16 (pinned host buffers, blocking), 2 concurrent callers per device as in predictors/nn.rs:64-67",
                "single_caller": world * BATCH * args.steps / e2e_single_s,
                "three_callers": world * BATCH * args.steps / e2e3_s},
        "e2e_packed": {"value": world * BATCH * args.steps / e2e_packed_s, "unit": UNIT,
                       "h2d_bytes_per_step": BATCH * nn.PACKED_DTYPE.itemsize, "d2h_bytes_per_step": BATCH * 363 * 2,
                       "call": "dg_engine_forward_packed"},
        "gpu_launches": launches * args.steps,
        "roofline": {"bound": "tensor", "kernel": "tower_kernel (persistent; 18 residual 3x3 128->128 convolutions per launch)",
                     "achieved": achieved_tf, "peak": peak_tf, "peak_kind": f"{peak_kind} burst bf16 (MEASURED_PEAKS.json)",
                     "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                     "us_per_launch": tower_launch_s * 1e6, "us_per_conv_layer": tower_launch_s * 1e6 / (2 * NUM_BLOCKS),
                     "whole_net_frac": (value_evals / world) * FLOP_PER_EVAL / (peak_tf * 1e12)},
    })
    if sustained is not None:
        tf = 2 * NUM_BLOCKS * BATCH * TOWER_CONV_FLOP_PER_POS / (sustained["tower_us_per_launch"] * 1e-6) / 1e12
        sustained["tower_tflops"] = tf
        if peak_sus:
            sustained["tower_frac_of_sustained_peak"] = tf / float(peak_sus)
            sustained["sustained_peak"] = float(peak_sus)
        line["sustained"] = sustained
    sp_summary = None
    if self_play is not None:
        try:
            sp_summary = self_play_summary(self_play)
        except Exception as exc:   # noqa: BLE001  (a summary must never cost the line)
            sp_summary = {"unavailable": repr(exc)[:200]}
    if sp_summary is not None:
        line["e2e"]["self_play"] = sp_summary
    if world == 1 and not os.environ.get("DG_BENCH_SKIP_CPU"):      # skipped only under the profiler
        line["cudnn_baseline"] = cudnn_baseline(tensors, feats, args.steps)
        if "value" in line["cudnn_baseline"]:      # the ratio the north star asks for: this engine vs the reference's cuDNN path, same box
            line["vs_cudnn"] = {"resident": value_evals / line["cudnn_baseline"]["value"], "e2e": e2e_evals / line["cudnn_baseline"]["e2e"]}
        rate, info = oracle_rate(seconds_target=15.0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
        line["host_features"] = host_feature_rates()
    if sp_summary is not None:
        line["self_play_summary"] = sp_summary          # last: the tail of stdout is what the driver's record keeps verbatim
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--self-play-seconds", type=float, default=0.0 if os.environ.get("DG_BENCH_SKIP_CPU") else 10.0,
                    help="length of each self-play sample appended to the line (0 = skip)")
    ap.add_argument("--sustained-seconds", type=float, default=0.0 if os.environ.get("DG_BENCH_SKIP_CPU") else 2.0,
                    help="device time of the sustained resident measurement (0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
