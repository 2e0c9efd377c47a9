#!/bin/bash
# round 2, call 10: full GPU suite after the fixes, ground truth of the largest default-vs-fp32
# differences, sanitizer runs of every device path, C2 launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert" | tail -8 > gpurun_out/r2c10_tests.txt
cat gpurun_out/r2c10_tests.txt
timeout 600 python profiles/probe_eig_truth.py 2>&1 | tail -1 > gpurun_out/r2c10_eig_truth.json
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c10_eig_truth.json"))
    print("truth: max err default %.2e  etol1e6 %.2e  fp32 %.2e  over %d etas" % (max(r["err_default"] for r in d), max(r["err_etol1e6"] for r in d), max(r["err_fp32"] for r in d), len(d)))
    for r in d[:4]: print(r)
except Exception as e: print("truth FAILED", e)
PY
timeout 900 compute-sanitizer --tool memcheck python profiles/sanitize_smoke.py 2>&1 | tail -6 > gpurun_out/r2_sanitizer_memcheck.txt; cat gpurun_out/r2_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python profiles/sanitize_smoke.py 2>&1 | tail -6 > gpurun_out/r2_sanitizer_racecheck.txt; cat gpurun_out/r2_sanitizer_racecheck.txt
cat > /tmp/c2.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
from scintools_b200 import BasicDyn, Dynspec
rng = np.random.default_rng(2)
dyn = rng.exponential(1.0, (4096, 8192)).astype(np.float32)
ds = Dynspec(dyn=BasicDyn(dyn, times=10.0*np.arange(8192), freqs=1400+0.03125*np.arange(4096), dt=10.0, df=0.03125), verbose=False)
for _ in range(3): ds.calc_sspec(dtype=np.float32)
for _ in range(3): ds.calc_acf(dtype=np.float32)
PY
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_c2_launches.csv python /tmp/c2.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_c2_launches.csv")) if len(r)>14 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=(r[0], r[4][:70]); agg.setdefault(k, {})[r[12]]=float(r[14])
last=list(agg.items())[-14:]
for (i,n),m in last: print(i, n, {k.split("__")[-1][:18]:v for k,v in m.items()})
PY
