#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_sparse_gpu.py -x -q -k "scene_graph" 2>&1 | tail -40
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -5
python bench.py --cpu-seconds 4 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -3 gpurun_out/r2h_bench.err; cut -c1-330 gpurun_out/r2h_bench.json
