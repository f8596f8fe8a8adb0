#!/bin/bash
# long samples (games reach the middle game, where a leaf costs the host more) with 4 cores per GPU: where should the priors / ladders be computed?
run() { label=$1; shift; line=$(env "$@" 2>>gpurun_out/r02_probe6.err | tail -1); echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.load(sys.stdin); print(round(d["value"],1), round(d["nn_evals_per_s"]), round(d["device_busy_frac"],2))')"; echo "{\"label\": \"$label\", \"line\": $line}" >> gpurun_out/r02_probe6.jsonl; }
: > gpurun_out/r02_probe6.jsonl
B="python tools/bench_selfplay.py --games 100000 --parallel 64 --seconds 150 --no-host-sample --threads 4"
run "4 cores, 64 games, 150 s, host priors + ladders" DG_X=1 taskset -c 0-3 $B --host-priors --host-ladders
run "4 cores, 64 games, 150 s, device priors" DG_X=1 taskset -c 0-3 $B --device-priors --host-ladders
run "4 cores, 64 games, 150 s, device priors + ladders" DG_X=1 taskset -c 0-3 $B --device-priors --device-ladders
