#!/bin/bash
# The 8-GPU measurements of round 2 (one box, 8 ranks, one engine per rank): the bench line, BASELINE configs[3] (64 concurrent
# games per GPU, 512 in total) and configs[4] (--ex-it extract) as fixed-duration samples.
#   tools/r02_n8.sh <ngpus> <cfg3 seconds> <cfg4 seconds>
n=${1:-8}; s3=${2:-120}; s4=${3:-60}
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
nproc > gpurun_out/r02_n${n}_nproc.txt
$run --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/r02_bench_n${n}.json 2> gpurun_out/r02_bench_n${n}.err
echo "bench rc=$?"
$run --master-port 29512 tools/bench_selfplay.py --games 25000 --parallel 64 --seconds $s3 --no-host-sample > gpurun_out/r02_selfplay_config3_n${n}.json 2> gpurun_out/r02_selfplay_config3_n${n}.err
echo "config3 rc=$?"
$run --master-port 29513 tools/bench_selfplay.py --games 1000000 --parallel 128 --rollouts 1 --ex-it --ex-it-rollouts 800 --seconds $s4 --no-host-sample > gpurun_out/r02_selfplay_config4_n${n}.json 2> gpurun_out/r02_selfplay_config4_n${n}.err
echo "config4 rc=$?"
tail -c 1200 gpurun_out/r02_selfplay_config3_n${n}.json; echo; tail -c 1200 gpurun_out/r02_selfplay_config4_n${n}.json
