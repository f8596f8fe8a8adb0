#!/bin/bash
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_selfplay_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02e_pytest.log; tail -3 gpurun_out/r02e_pytest.log
python tools/tower_trace.py 256 0 > gpurun_out/r02e_trace.txt 2>&1; cp gpurun_out/tower_trace.npy gpurun_out/r02e_trace.npy; head -4 gpurun_out/r02e_trace.txt
python tools/ab_resident.py build/libdg_engine_r01.so dream_go_b200/libdg_engine.so > gpurun_out/r02e_ab.json 2> gpurun_out/r02e_ab.err; cat gpurun_out/r02e_ab.json
python tools/bench_selfplay.py --games 100000 --parallel 128 --seconds 8 --no-host-sample 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.load(sys.stdin); print("selfplay 128:", round(d["value"],1), round(d["nn_evals_per_s"]))'
