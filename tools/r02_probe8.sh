#!/bin/bash
timeout 200 python -m pytest tests/test_selfplay_gpu.py -m gpu -x -q --timeout 100 --timeout-method thread > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02n_pytest.log
run() { label=$1; shift; line=$(env "$@" 2>>gpurun_out/r02_probe8.err | tail -1); echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.load(sys.stdin); print(round(d["value"],1), round(d["nn_evals_per_s"]), round(d["device_busy_frac"],2))')"; echo "{\"label\": \"$label\", \"line\": $line}" >> gpurun_out/r02_probe8.jsonl; }
: > gpurun_out/r02_probe8.jsonl; : > gpurun_out/r02_probe8.err
B="python tools/bench_selfplay.py --games 100000 --no-host-sample"
run "16 threads, 32 games, 8 s, auto" DG_SELFPLAY_TRACE=1 $B --parallel 32 --seconds 8
run "16 threads, 128 games, 8 s, auto" DG_SELFPLAY_TRACE=1 $B --parallel 128 --seconds 8
run "4 cores, 64 games, 70 s, auto" DG_SELFPLAY_TRACE=1 taskset -c 0-3 $B --parallel 64 --seconds 70 --threads 4
grep "priors on the device" gpurun_out/r02_probe8.err
