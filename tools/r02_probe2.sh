#!/bin/bash
python tools/r02_probe2.py > gpurun_out/r02_batch_scaling.json 2> gpurun_out/r02_batch_scaling.err; tail -c 1500 gpurun_out/r02_batch_scaling.json; echo
run() { label=$1; shift; line=$(env "$@" 2>>gpurun_out/r02_probe2.err | tail -1); echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.load(sys.stdin); print(round(d["value"],1), round(d["nn_evals_per_s"]), round(d["mean_batch"],1))')"; echo "{\"label\": \"$label\", \"line\": $line}" >> gpurun_out/r02_probe2.jsonl; }
: > gpurun_out/r02_probe2.jsonl
B="python tools/bench_selfplay.py --games 100000 --seconds 8 --no-host-sample"
run "128 games, 4 groups (256/batch)" DG_X=1 $B --parallel 128
run "128 games, 2 groups (512/batch)" DG_SELFPLAY_GROUPS=2 $B --parallel 128
run "256 games, 4 groups (512/batch)" DG_SELFPLAY_GROUPS=4 $B --parallel 256
run "192 games, 3 groups (512/batch)" DG_SELFPLAY_GROUPS=3 $B --parallel 192
run "4 cores, 128 games, 2 groups, host priors" DG_SELFPLAY_GROUPS=2 taskset -c 0-3 $B --parallel 128 --threads 4 --host-priors
run "4 cores, 256 games, 4 groups, host priors" DG_SELFPLAY_GROUPS=4 taskset -c 0-3 $B --parallel 256 --threads 4 --host-priors
run "4 cores, 128 games, 4 groups, host priors" DG_X=1 taskset -c 0-3 $B --parallel 128 --threads 4 --host-priors
