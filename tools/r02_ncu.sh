#!/bin/bash
# Round-2 profiler captures (one GPU): launch list of the bench step, full capture of the tower kernel, the small kernels of the
# resident step and of the raw-position path (feature kernel with the device ladder reader, prior kernel, compact pack).
export DG_BENCH_SKIP_CPU=1
B="python bench.py --steps 3 --warmup 3 --self-play-seconds 0 --sustained-seconds 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench_steps3.csv $B > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tower_kernel -s 4 -c 1 -f -o gpurun_out/r02_tower $B > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k "regex:pack_|policy_fc|heads_finish" -s 8 -c 6 -f -o gpurun_out/r02_small $B > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k "regex:planes_from_stones|prior_from_policy|pack_compact" -s 4 -c 6 -f -o gpurun_out/r02_raw python tools/profile_raw_path.py > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_launches_raw_path.csv python tools/profile_raw_path.py > /dev/null 2>&1
ls -la gpurun_out/r02_*.ncu-rep gpurun_out/r02_launches_*.csv
