"""The queue-driven self-play driver (dg_selfplay_run_engine: polling scheduler, round-robin over engines, deadline handling)
without a device: tools/host_sanitize.cpp links the host half of the path against a functional CPU stand-in for the engine's
leaf-batch queue (legal moves from raw stones and hashes, evaluated by a "device" thread per batch) and checks that the games do
not depend on workers, groups, engines or on where the priors are built, that a deadline drops the batches in flight cleanly, and that the stand-in's legal
masks are the board's over a 600-ply game.  (The same program is what runs under ASan / UBSan / TSan:
profiles/r02_host_asan_ubsan.log, r02_host_tsan.log.)"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_queue_driver_on_the_cpu_stand_in(tmp_path):
    exe = tmp_path / "host_queue"
    cmd = ["g++", "-O1", "-march=x86-64-v3", "-ffp-contract=off", "-std=c++17", "-I" + os.path.join(ROOT, "dream_go_b200", "csrc"),
           os.path.join(ROOT, "tools", "host_sanitize.cpp"), os.path.join(ROOT, "dream_go_b200", "csrc", "search_api.cpp"),
           os.path.join(ROOT, "dream_go_b200", "csrc", "go_api.cpp"), "-o", str(exe), "-lpthread"]
    subprocess.check_call(cmd, cwd=ROOT)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[-1] == "board api ok" and "FAILED" not in out.stdout
    queue = [l for l in lines if l.startswith("queue variant")]
    assert len(queue) == 6 and all("games 7 moves 168" in l and "still held 0" in l for l in queue)
    assert len({l.split("evals")[1].split()[0] for l in queue}) == 1          # the same evaluations whatever the schedule
    table = [l for l in lines if l.startswith("queue shared-table")]            # one process-wide table under the queue-driven driver
    assert len(table) == 2 and all("games 8 moves 64" in l and "held 0" in l for l in table)
    # random configurations (games, rollouts, probes, threads, groups, one or two engines, where the priors are built, --ex-it,
    # per-game tables): every run ends, gives its leaf batches back and plays the games of the plainest schedule
    fuzz = subprocess.run([str(exe), "fuzz", "80", "3"], capture_output=True, text=True, timeout=600)
    assert fuzz.returncode == 0 and "80 cases" in fuzz.stdout and " 0 failures" in fuzz.stdout, fuzz.stdout[-2000:]
