#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_vote_gpu.py tests/test_sparse_gpu.py -x -q 2>&1 | tail -6
python -m pytest tests/test_parity_gpu.py -q -s > gpurun_out/r2f_parity_tests.log 2>&1
grep -E "engine C2|class_pred|boxes|bn-|passed|failed|Error" gpurun_out/r2f_parity_tests.log
python tools/time_vote.py > gpurun_out/r2f_time_vote.txt 2>&1; cat gpurun_out/r2f_time_vote.txt
python bench.py --steps 20 --train-steps 2 --cpu-seconds 2 > gpurun_out/r2f_bench_short.json 2> gpurun_out/r2f_bench_short.err
tail -5 gpurun_out/r2f_bench_short.err; cut -c1-1200 gpurun_out/r2f_bench_short.json
