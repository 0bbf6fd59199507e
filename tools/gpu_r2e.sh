#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'hv_bin|hv_plan|hv_tile' -s 8 -c 4 -o gpurun_out/r2e_vote -f python tools/time_vote.py > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/r2e_vote.ncu-rep
