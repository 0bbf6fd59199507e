#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_sparse_gpu.py -x -q -k "scene_graph" 2>&1 | tail -15
python bench.py --cpu-seconds 2 --train-steps 0 > gpurun_out/r2i_bench_prio.json 2> gpurun_out/r2i_bench_prio.err; tail -3 gpurun_out/r2i_bench_prio.err; cut -c1-330 gpurun_out/r2i_bench_prio.json
CVB200_GRAPH_NO_PRIORITY=1 python bench.py --cpu-seconds 2 --train-steps 0 > gpurun_out/r2i_bench_noprio.json 2> gpurun_out/r2i_bench_noprio.err; cut -c1-330 gpurun_out/r2i_bench_noprio.json
python bench.py --cpu-seconds 2 --train-steps 0 --streams 2 2>/dev/null | cut -c1-330
python bench.py --cpu-seconds 2 --train-steps 0 --streams 4 2>/dev/null | cut -c1-330
