#!/bin/bash
# round 2, GPU session 11: image kernels with one AGM ladder per ray + pow special cases; jet timing; tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/s11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s11_pytest.log
tail -4 gpurun_out/s11_pytest.log
timeout 600 python tools/image_bench.py > gpurun_out/s11_image.txt 2>&1
cat gpurun_out/s11_image.txt
timeout 600 python tools/jet_bench.py > gpurun_out/s11_jet.json 2> gpurun_out/s11_jet.err; cat gpurun_out/s11_jet.json; tail -3 gpurun_out/s11_jet.err
KGEO_JET_NO_CPU=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_jet_launches.csv python tools/jet_bench.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_jet_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=collections.defaultdict(float); cnt=collections.Counter()
for r in rows[1:]:
    v=float(r[vi].replace(',','')); v*= {'ns':1e-6,'us':1e-3,'ms':1.0,'s':1e3}.get(r[ui],1e-6)
    k=r[ki].split('(')[0][:40]; tot[k]+=v; cnt[k]+=1
for k,v in sorted(tot.items(), key=lambda kv:-kv[1]): print('%-42s %5d launches %9.3f ms'%(k,cnt[k],v))
PY
