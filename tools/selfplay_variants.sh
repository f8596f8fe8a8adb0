#!/bin/bash
# Self-play throughput under scarce host cores (the 8-GPU box has 4 hardware threads per GPU): napping (default) vs spinning
# engine calls, device priors A/B.
#   tools/selfplay_variants.sh <out.jsonl> [seconds]
out=${1:-gpurun_out/selfplay_variants.jsonl}
secs=${2:-10}
: > "$out"
run() {   # label, cpu list ('' = all), threads, extra args
    label=$1; cpus=$2; threads=$3; shift 3
    pre=""
    [ -n "$cpus" ] && pre="taskset -c $cpus"
    line=$($pre python tools/bench_selfplay.py --games 100000 --parallel 128 --seconds "$secs" --threads "$threads" --no-host-sample "$@" 2>/dev/null | tail -1)
    echo "{\"label\": \"$label\", \"line\": $line}" >> "$out"
    echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.load(sys.stdin); print(round(d["value"],1), round(d["nn_evals_per_s"]), round(d["mean_batch"],1))')"
}
run "16 threads" "" 16
run "16 threads, spinning engine calls" "" 16 --spin-sync
run "4 cores" 0-3 4
run "4 cores, spinning engine calls" 0-3 4 --spin-sync
run "4 cores, device priors" 0-3 4 --device-priors
run "4 cores, 64 games" 0-3 4 --parallel 64
run "8 cores" 0-7 8
