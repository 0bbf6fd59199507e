#!/bin/bash
# round-2 first GPU call: baseline state + probes (everything lands in gpurun_out/)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r2a_tests.log 2>&1
tools/probes/_bin/red_probe 128 381529 > gpurun_out/r2a_red_probe.txt 2>&1
tools/probes/_bin/red_probe 256 2500000 >> gpurun_out/r2a_red_probe.txt 2>&1
python tools/conv_probe.py C2 all > gpurun_out/r2a_conv_probe.txt 2>&1
for lv in 16 8 4 1; do python tools/conv_trace.py 0 $lv > gpurun_out/r2a_conv_trace_L$lv.txt 2>&1; done
python bench.py --cpu-seconds 4 2>&1 | tail -1 > gpurun_out/r2a_bench.json
python tools/host_profile.py > gpurun_out/r2a_host_profile.txt 2>&1
tail -3 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_red_probe.txt; cat gpurun_out/r2a_conv_probe.txt
