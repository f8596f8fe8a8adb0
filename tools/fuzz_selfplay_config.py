#!/usr/bin/env python
"""Random configurations of the self-play driver (csrc/search_api.cpp) on the host-only predictor: every run must end (no hang
-- each case runs in its own process under a timeout), return 0, finish exactly the games it was asked for, give the same
digest for another number of worker threads / groups, and every record must replay legally on the oracle board.  CPU only.

    python tools/fuzz_selfplay_config.py [--cases 60] [--seed 1] [--out profiles/r02_fuzz_selfplay_config.log]
"""
import argparse
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, re, sys
sys.path.insert(0, %r)
from dream_go_b200 import mcts as pm
from oracle import go as ogo
cfg = json.loads(sys.argv[1])
alt = cfg.pop("alt")
st, games = pm.self_play(pm.RandomPredictor(), **cfg)
assert st["games_finished"] == cfg["num_games"] and len(games) == cfg["num_games"], (st, len(games))
ogo.use_default_zobrist()
for sgf in games:
    komi = float(re.search(r"KM\[([-0-9.]+)\]", sgf).group(1))
    board = ogo.Board(komi)
    plies = 0
    for m in re.finditer(r";([BW])\[([a-s]{0,2})\]", sgf):
        color = 1 if m.group(1) == "B" else 2
        if m.group(2):
            x, y = ord(m.group(2)[0]) - 97, ord(m.group(2)[1]) - 97
            assert board.is_valid(color, x, y), sgf[:200]
            board.place(color, x, y)
        plies += 1
    assert plies >= 1
deterministic = not (cfg.get("cache_shared", 0) and cfg.get("cache_capacity", 0))
if deterministic:
    st2, _ = pm.self_play(pm.RandomPredictor(), **{**cfg, **alt})
    assert st2["digest"] == st["digest"] and st2["moves"] == st["moves"], (st, st2)
print(json.dumps({"moves": int(st["moves"]), "evals": int(st["evals"]), "searches": int(st["searches"]), "hits": int(st["cache_hits"])}))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    rng = random.Random(args.seed)
    lines, bad = [], 0
    for k in range(args.cases):
        games = rng.choice([1, 2, 3, 5, 8, 13])
        cfg = dict(num_games=games, num_parallel=rng.choice([1, 2, 3, 4, 7, 16, 40]), num_rollout=rng.choice([1, 2, 9, 24, 60, 130]),
                   probes_per_round=rng.choice([1, 2, 3, 8, 12, 20]), max_plies=rng.choice([1, 2, 9, 30, 70]),
                   num_threads=rng.choice([1, 2, 3, 8]), seed=rng.randrange(1, 1 << 30), num_groups=rng.choice([0, 1, 2, 3, 5, 8, 11]),
                   ex_it=rng.random() < 0.3, num_ex_it_rollout=rng.choice([5, 40, 90]), cache_capacity=rng.choice([0, 0, 1, 7, 300, 5000]),
                   cache_shared=rng.choice([0, 0, 1, 4, 64]), temperature=rng.choice([0.8, 0.05, 2.0]), dirichlet_noise=rng.choice([0.25, 0.0, 1.0]))
        cfg["alt"] = dict(num_threads=rng.choice([1, 2, 5]), num_groups=rng.choice([0, 1, 2, 4, 8]))
        try:
            out = subprocess.run([sys.executable, "-c", CHILD % ROOT, json.dumps(cfg)], capture_output=True, text=True, timeout=300, cwd=ROOT)
            ok = out.returncode == 0
            tail = out.stdout.strip().splitlines()[-1] if ok and out.stdout.strip() else (out.stderr.strip().splitlines() or ["?"])[-1][:300]
        except subprocess.TimeoutExpired:
            ok, tail = False, "TIMEOUT (hang)"
        bad += not ok
        line = f"case {k:3d} {'ok  ' if ok else 'FAIL'} {json.dumps(cfg, sort_keys=True)} -> {tail}"
        print(line, flush=True)
        lines.append(line)
    lines.append(f"{args.cases} cases, {bad} failures")
    print(lines[-1])
    if args.out:
        with open(args.out, "w") as fh:
            fh.write("\n".join(lines) + "\n")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
