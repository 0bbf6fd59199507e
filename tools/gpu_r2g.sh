#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_sparse_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -4
# launch list of scenes 2 and 3 (time + DRAM bytes per launch)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_scene_C2.csv python tools/profile_scene.py C2 3 > gpurun_out/r2g_ncu_a.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2g_launches_scene_C2.csv")) if len(r)>10]
h=rows[0]; k=h.index("Kernel Name"); v=h.index("Metric Value"); m=h.index("Metric Name"); i=h.index("ID")
per=collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[i],{"name":r[k].split("(")[0][:44]})[r[m]]=float(r[v].replace(",",""))
ids=list(per)
n=len(ids)//3
last=ids[2*n:]
agg=collections.OrderedDict()
for idn in last:
    d=per[idn]; a=agg.setdefault(d["name"],[0,0.0,0.0]); a[0]+=1; a[1]+=d.get("gpu__time_duration.sum",0)/1e3; a[2]+=d.get("dram__bytes_read.sum",0)+d.get("dram__bytes_write.sum",0)
print("launches per scene", n)
for nm,a in agg.items(): print("%-46s %4d launches %9.1f us %9.1f MB dram"%(nm,a[0],a[1],a[2]/1e6))
PY
# full captures: the six last convolutions (block8 + final) of scene 2, and the vote kernels
ncu --set full --clock-control none --import-source on -k regex:'sc_conv_persist' -s 114 -c 12 -o gpurun_out/r2g_conv -f python tools/profile_scene.py C2 2 > gpurun_out/r2g_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hv_scatter|hv_finalize' -s 2 -c 2 -o gpurun_out/r2g_vote -f python tools/profile_scene.py C2 2 > gpurun_out/r2g_ncu_c.log 2>&1
ls -la gpurun_out/r2g_*.ncu-rep
python bench.py --cpu-seconds 6 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -3 gpurun_out/r2g_bench.err; cut -c1-400 gpurun_out/r2g_bench.json
