#!/bin/bash
# Self-play throughput: leaf-batch queue driver vs the blocking-call driver, under plenty of and scarce host cores (the 8-GPU
# box has 4 hardware threads per GPU), host vs device priors.
#   tools/selfplay_variants.sh <out.jsonl> [seconds]
out=${1:-gpurun_out/selfplay_variants.jsonl}
secs=${2:-10}
: > "$out"
run() {   # label, cpu list ('' = all), threads, extra args
    label=$1; cpus=$2; threads=$3; shift 3
    pre=""
    [ -n "$cpus" ] && pre="taskset -c $cpus"
    line=$(DG_SELFPLAY_TRACE=${TRACE:-} $pre python tools/bench_selfplay.py --games 100000 --parallel 128 --seconds "$secs" --threads "$threads" --no-host-sample "$@" 2>>"$out.err" | tail -1)
    echo "{\"label\": \"$label\", \"line\": $line}" >> "$out"
    echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.load(sys.stdin); print(round(d["value"],1), round(d["nn_evals_per_s"]), round(d["mean_batch"],1))')"
}
run "all threads, queue" "" 0
run "all threads, blocking calls" "" 0 --blocking-calls
run "all threads, queue, no graph" "" 0 --no-graph
run "4 cores, queue (device priors)" 0-3 4
run "4 cores, queue, host priors" 0-3 4 --host-priors
run "4 cores, blocking calls" 0-3 4 --blocking-calls
run "4 cores, blocking calls, device priors" 0-3 4 --blocking-calls --device-priors
run "32 games, all threads, queue" "" 0 --parallel 32
run "32 games, all threads, blocking calls" "" 0 --parallel 32 --blocking-calls
run "64 games, 4 cores, queue" 0-3 4 --parallel 64
