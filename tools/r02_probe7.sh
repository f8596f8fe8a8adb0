#!/bin/bash
timeout 200 python -m pytest tests/test_selfplay_gpu.py -m gpu -x -q --timeout 100 --timeout-method thread > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02m_pytest.log
run() { label=$1; shift; line=$(env "$@" 2>>gpurun_out/r02_probe7.err | tail -1); echo "$label: $(echo "$line" | python -c 'import sys,json; d=json.load(sys.stdin); print(round(d["value"],1), round(d["nn_evals_per_s"]), round(d["device_busy_frac"],2))')"; echo "{\"label\": \"$label\", \"line\": $line}" >> gpurun_out/r02_probe7.jsonl; }
: > gpurun_out/r02_probe7.jsonl; : > gpurun_out/r02_probe7.err
B="python tools/bench_selfplay.py --games 100000 --no-host-sample"
run "16 threads, 32 games, 10 s, auto" DG_SELFPLAY_TRACE=1 $B --parallel 32 --seconds 10
run "16 threads, 32 games, 10 s, host priors" DG_X=1 $B --parallel 32 --seconds 10 --host-priors
run "16 threads, 128 games, 10 s, auto" DG_SELFPLAY_TRACE=1 $B --parallel 128 --seconds 10
run "16 threads, 128 games, 10 s, host priors" DG_X=1 $B --parallel 128 --seconds 10 --host-priors
run "4 cores, 128 games, 10 s, auto" DG_SELFPLAY_TRACE=1 taskset -c 0-3 $B --parallel 128 --seconds 10 --threads 4
run "4 cores, 64 games, 90 s, auto" DG_SELFPLAY_TRACE=1 taskset -c 0-3 $B --parallel 64 --seconds 90 --threads 4
grep "priors on the device" gpurun_out/r02_probe7.err
