#!/bin/bash
# round 2, call 13 (call 12 re-run after the container was replaced; full bench leg moved to the next call): tensor-core mat-vec of the default eigen solver (mma.sync on the block
# layout) against the packed-FMA mat-vec: A/B bench, error against the fp32 solver on all
# 1024 curvatures, GPU tests, ncu capture, racecheck of both solvers
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], d['sweep']['iters_hist'], 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
SB_EIG_NO_TC=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-extra 2>gpurun_out/r2c13_notc.err | tail -1 > gpurun_out/r2c13_bench_notc.json
show "bench NO_TC" gpurun_out/r2c13_bench_notc.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-extra 2>gpurun_out/r2c13_tc.err | tail -1 > gpurun_out/r2c13_bench_tc.json
show "bench TC" gpurun_out/r2c13_bench_tc.json
tail -3 gpurun_out/r2c13_tc.err
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c13_eig_error.json; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c13_eig_error.json"))
    print("fp32 iters", d["fp32_iters_mean"])
    for k in ("default","no_tc","rtol5e4","rtol3e3","etol5e7"):
        print(k, "vs fp32: max %.2e p99 %.2e iters %.2f gt24 %d" % (d[k]["max"], d[k]["p99"], d[k]["iters_mean"], d[k]["iters_gt24"]), d[k]["worst"][:2])
except Exception as ex:
    print("eig_error FAILED", ex)
PY
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert|FAILED" | tail -8 > gpurun_out/r2c13_tests.txt
cat gpurun_out/r2c13_tests.txt
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build|row_fft_r2c|tile_fft_tma" -s 12 -c 5 \
    -o gpurun_out/r2c13_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c13_ncu.log 2>&1
cat > /tmp/race.py <<'PY'
import os, sys
sys.path.insert(0, ".")
import numpy as np
from scintools_b200 import ththmod as thth
rng = np.random.default_rng(0)
nf, nt, npad = 32, 64, 1
d0 = rng.normal(size=(nf, nt)); d0 -= d0.mean()
t = np.arange(nt) * 10.0; f = 1400.0 + np.arange(nf) * 0.05
fd = thth.fft_axis(t, "mHz", npad); tau = thth.fft_axis(f, "us", npad)
edges = np.linspace(-20, 20, 96); etas = np.linspace(0.002, 0.02, 4)
cs = thth.conjugate_spectrum(d0, npad, 0.0)
print("tensor-core", thth.eta_sweep(cs, tau, fd, etas, edges))
os.environ["SB_EIG_NO_TC"] = "1"
print("packed-FMA ", thth.eta_sweep(cs, tau, fd, etas, edges))
os.environ["SB_EIG_FP32"] = "1"
print("fp32       ", thth.eta_sweep(cs, tau, fd, etas, edges))
PY
timeout 420 compute-sanitizer --tool racecheck --print-limit 3 python /tmp/race.py 2>&1 | tail -30 > gpurun_out/r2_sanitizer_racecheck.txt; cut -c1-260 gpurun_out/r2_sanitizer_racecheck.txt | tail -14
timeout 240 compute-sanitizer --tool memcheck --print-limit 3 python /tmp/race.py 2>&1 | tail -8 > gpurun_out/r2_sanitizer_memcheck_tc.txt; cut -c1-200 gpurun_out/r2_sanitizer_memcheck_tc.txt | tail -5
