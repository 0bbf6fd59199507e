#!/bin/bash
# tools/gpu_check.sh TAG [pytest-args] -- runs ON THE GPU BOX (via gpurun): GPU tests, the C2/C5 vote benches,
# the ncu launch list of the bench and one --set full capture of our kernels; everything lands in gpurun_out/.
TAG=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -12
python bench.py --cpu-seconds 6 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}.json | cut -c1-1500
python bench.py --workload C5 --steps 10 --cpu-seconds 4 2>&1 | tail -1 | tee gpurun_out/bench_${TAG}_c5.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hv_|sc_|bp_|head_' -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/ncu_b.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_${TAG}.csv")) if len(r)>10]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value")
import collections
agg=collections.OrderedDict()
for r in rows[1:]:
    nm=r[k].split('(')[0][:60]; agg.setdefault(nm,[0,0.0]); agg[nm][0]+=1; agg[nm][1]+=float(r[v])
tot=sum(a[1] for a in agg.values())
for nm,(c,t) in agg.items(): print('%-62s %5d launches %10.1f us %5.1f%%'%(nm,c,t/1e3,100*t/tot))
PY
ncu --set full --clock-control none --import-source on -k regex:'sc_conv_tc|hv_scatter|hv_finalize' -s 60 -c 12 -o gpurun_out/prof_${TAG} -f python bench.py --steps 3 --warmup 3 --cpu-seconds 1 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | tail -5
