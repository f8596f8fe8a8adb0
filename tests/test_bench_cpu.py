"""Host-side pieces of bench.py that can run without a device: the compact self-play summary (the part of the line that
survives the driver's parsing) on the recorded round-2 lines, the reference arm's contract keys, the roofline inputs."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def recorded_line(name):
    line = None
    with open(os.path.join(ROOT, "profiles", name)) as fh:
        for text in fh:
            text = text.strip()
            if text.startswith("{"):
                line = json.loads(text)
    return line


def test_self_play_summary_of_the_recorded_lines():
    import bench
    for name, gpus in (("r02_bench_n1.json", 1), ("r02_bench_n8.json", 8)):
        line = recorded_line(name)
        assert line["n_gpus"] == gpus and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
        summary = bench.self_play_summary(line["self_play"])
        json.dumps(summary)                                     # plain numbers only
        for key in ("configs2_32_games_per_gpu", "configs3_shape_64_games_per_gpu", "games128_per_gpu"):
            moves, evals = summary[key]
            assert moves > 0 and evals > 100 * moves            # ~800 evaluations per move
        if gpus == 1:
            assert summary["configs2_vs_cudnn_reference_batch16"] > 1.0
    # a sample that is missing costs the summary, never the line (run_ours wraps the call)
    try:
        bench.self_play_summary({"host_threads_per_gpu": 4})
    except KeyError:
        pass


def test_roofline_constants_follow_the_survey():
    import bench
    # 976,681,890 MAC per evaluation, 53.23 M MAC per 3x3 128->128 convolution and position
    assert bench.FLOP_PER_EVAL == 2 * 976_681_890
    assert bench.TOWER_CONV_FLOP_PER_POS == 2 * 361 * 9 * 128 * 128
    peak_tf, peak_gbs, kind = bench.measured_peaks()
    assert peak_tf > 100 and peak_gbs > 1000 and kind in ("measured", "fallback")
    cfg = bench.workload_config(8)
    assert "configs[1]" in cfg["workload"] and cfg["batch_per_gpu"] == 256 and "model" not in cfg


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no device: the CPU port of dg_nn::forward on a bounded sample."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "nn_evals_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("residual tower 9-block x 128-filter")
    import bench
    assert line["config"] == bench.workload_config(1)            # both arms print the same config dictionary


def test_self_play_front_end_prints_one_record_per_game():
    """tools/self_play.py (`dream_go --self-play N`): records on stdout, progress on stderr; host-only predictor here."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "self_play.py"), "--self-play", "3", "--num-games", "2", "--num-rollout", "12",
                          "--ex-it", "--num-ex-it-rollout", "20", "--host-only", "--seed", "5", "--num-threads", "2"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    records = out.stdout.strip().splitlines()
    assert len(records) == 3 and all(r.startswith("(;GM[1]FF[4]SZ[19]RU[Chinese]KM[") and r.endswith(")") for r in records)
    assert out.stderr.startswith("...") and "3 games" in out.stderr
    # and without a device the engine path fails loudly instead of falling back
    bad = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "self_play.py"), "--self-play", "1", "--random-weights"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    import torch
    if not torch.cuda.is_available():
        assert bad.returncode != 0 and "Cuda" in bad.stderr


def test_line_assembly_of_the_device_arm_with_the_device_mocked(monkeypatch, capsys):
    """bench.run_ours end to end with every device-facing piece replaced (engine, shards, clock sampler, cuDNN and oracle legs):
    the line carries every key of the contract at N = 1 and N = 8, the self-play summary sits inside `e2e` and at the end."""
    import types
    import numpy as np
    import bench
    from dream_go_b200 import nn, shard
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_selfplay

    class FakeNet:
        def __init__(self, *a, **k):
            self.max_batch = k.get("max_batch", 256)

        @classmethod
        def from_tensors(cls, tensors, **k):
            return cls(**k)

        def pinned(self, shape, dtype):
            return np.zeros(shape, dtype)

        def forward_into(self, feats, value, policy, packed=False):
            policy[...] = np.float16(1.0 / 362)

        def synchronize(self):
            pass

        def time_resident(self, batch, steps, tower=False, flush_l2=False):
            return 0.4 * steps, 0.36 * steps, 4

        def time_e2e(self, feats, steps, callers=2):
            return 0.0004 * steps

        def close(self):
            pass

    class FakeShards:
        def __init__(self, backend="nccl"):
            pass

        def barrier(self):
            pass

        def max(self, x):
            return x

        def sum(self, x):
            return x

        def close(self):
            pass

        def selfplay_totals(self, st):
            return {"moves_per_s": 500.0, "nn_evals_per_s": 400000.0, "mean_device_batch": 250.0, "predictor_seconds": 9.0, "seconds": 10.0}

    class FakeClock:
        def __init__(self, device):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def summary(self):
            return {"sm_mhz": 1900.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 10}

    monkeypatch.setattr(nn, "Network", FakeNet)
    monkeypatch.setattr(nn, "pack_positions", lambda feats: np.zeros(feats.shape[0], nn.PACKED_DTYPE))
    monkeypatch.setattr(shard, "Shards", FakeShards)
    monkeypatch.setattr(bench_selfplay, "sample", lambda net, **k: ({"cache_hits": 5, "evals": 100, "moves": 10, "seconds": 10.0}, []))
    monkeypatch.setattr(bench, "ClockSampler", FakeClock)
    monkeypatch.setattr(bench, "cudnn_baseline", lambda tensors, feats, steps: {"value": 260000.0, "unit": bench.UNIT, "ms_per_step": 0.97, "e2e": 180000.0})
    monkeypatch.setattr(bench, "oracle_rate", lambda seconds_target, steps=1, warmup=0: (140.0, {"cores": 16, "kind": "port", "seconds": 1.8, "sample": "mock"}))
    monkeypatch.setattr(bench, "host_feature_rates", lambda positions=512: {"unit": "positions/s"})
    args = types.SimpleNamespace(gpus=1, steps=20, warmup=3, self_play_seconds=0.5, sustained_seconds=0.5, impl="ours")
    contract = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "clocks", "e2e", "gpu_launches", "roofline"}
    for world in (1, 8):
        bench.run_ours(args, 0, world, 0)
        line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
        assert contract <= set(line) and line["n_gpus"] == world and line["config"] == bench.workload_config(world)
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"]) and 0 < line["roofline"]["frac"] < 1.2
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "self_play"} <= set(line["e2e"])
        assert list(line)[-1] == "self_play_summary" and line["self_play_summary"] == line["e2e"]["self_play"]
        assert ("cpu_baseline" in line) == (world == 1) and ("vs_cudnn" in line) == (world == 1)
        assert line["gpu_launches"] == 4 * args.steps and line["vs_baseline"] is None
