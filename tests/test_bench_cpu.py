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
