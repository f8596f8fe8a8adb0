"""The N > 1 path on the CPU: two ranks over gloo, each with its own shard of self-play games (host-only predictor),
counters added up exactly as bench.py / tools/bench_selfplay.py do under NCCL.  No data-path collective exists to test --
what is checked is that shards are independent (different games), reproducible, and that the totals are the sum."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, %r)
from dream_go_b200 import mcts, shard
sh = shard.Shards(backend="gloo")
st, games = mcts.self_play(mcts.RandomPredictor(), num_games=3, num_parallel=3, num_rollout=20, probes_per_round=2, max_plies=10,
                           num_threads=1, seed=sh.seed(100))
sh.barrier()
tot = sh.selfplay_totals(st)
print(json.dumps({"rank": sh.rank, "world": sh.world, "digest": st["digest"], "moves": st["moves"], "evals": st["evals"], "totals": tot}))
sh.close()
''' % ROOT


def run(world: int, port: int):
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    out = []
    for p in procs:
        stdout, stderr = p.communicate(timeout=300)
        assert p.returncode == 0, stderr[-2000:]
        out.append(json.loads(stdout.strip().splitlines()[-1]))
    return sorted(out, key=lambda r: r["rank"])


def test_two_shards_over_gloo():
    two = run(2, 29631)
    assert [r["rank"] for r in two] == [0, 1] and all(r["world"] == 2 for r in two)
    assert two[0]["digest"] != two[1]["digest"]                       # every shard plays its own games
    tot = two[0]["totals"]
    assert tot == two[1]["totals"] or abs(tot["seconds"] - two[1]["totals"]["seconds"]) < 1e-9
    assert tot["moves"] == two[0]["moves"] + two[1]["moves"] == 60     # counters add up: 2 shards x 3 games x 10 plies
    assert tot["evals"] == two[0]["evals"] + two[1]["evals"]
    one = run(1, 29632)                                                # rank 0's shard alone plays the same games
    assert one[0]["digest"] == two[0]["digest"] and one[0]["totals"]["moves"] == 30
