#!/bin/bash
run() { label=$1; shift; line=$(env "$@" 2>>gpurun_out/r02_probe5.err | tail -1); echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.load(sys.stdin); print(round(d["value"],1), round(d["nn_evals_per_s"]), round(d["mean_batch"],1))')"; echo "{\"label\": \"$label\", \"line\": $line}" >> gpurun_out/r02_probe5.jsonl; }
: > gpurun_out/r02_probe5.jsonl
B="python tools/bench_selfplay.py --games 100000 --seconds 8 --no-host-sample"
run "32 games, 2 groups (default)" DG_X=1 $B --parallel 32
run "32 games, 1 group" DG_SELFPLAY_GROUPS=1 $B --parallel 32
run "32 games, 3 groups" DG_SELFPLAY_GROUPS=3 $B --parallel 32
run "32 games, 2 groups, 16 probes" DG_X=1 $B --parallel 32 --probes 16
run "64 games, 2 groups (default)" DG_X=1 $B --parallel 64
run "64 games, 3 groups" DG_SELFPLAY_GROUPS=3 $B --parallel 64
run "128 games, device ladders" DG_X=1 $B --parallel 128 --device-ladders
run "128 games, shared cache 200k" DG_X=1 $B --parallel 128 --cache 200000 --shared-cache 64
run "4 cores, 128 games, device ladders + priors" DG_X=1 taskset -c 0-3 $B --parallel 128 --threads 4 --device-ladders --device-priors
run "2 cores, 128 games, host all" DG_X=1 taskset -c 0-1 $B --parallel 128 --threads 2 --host-priors --host-ladders
run "2 cores, 128 games, device priors" DG_X=1 taskset -c 0-1 $B --parallel 128 --threads 2 --device-priors --host-ladders
run "2 cores, 128 games, device priors + ladders" DG_X=1 taskset -c 0-1 $B --parallel 128 --threads 2 --device-priors --device-ladders
