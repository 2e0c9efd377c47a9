#!/bin/bash
# round 2, call 11: validation after the Newton-polished check / etol_h / pinned pool;
# racecheck with full output; the CPU reference arm twice on this host
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert" | tail -8 > gpurun_out/r2c11_tests.txt
cat gpurun_out/r2c11_tests.txt
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/r2c11_bench.err | tail -1 > gpurun_out/r2c11_bench.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c11_bench.json").read())
x=d.get('extra') or {}; c=d.get('cpu_baseline') or {}
print("bench", round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], d['sweep']['iters_hist'], 'e2e', round(d['e2e']['value']), 'e2e_f64', d['e2e_f64'] and round(d['e2e_f64']['value']), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3)) for k,v in x.items()}, 'cpu', c.get('value'), 'err', c.get('max_rel_err_vs_gpu'), 'strong', d.get('strong') and round(d['strong']['value']), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'])
PY
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c11_eig_error.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c11_eig_error.json"))
print("fp32 iters", d["fp32_iters_mean"], "default vs fp32: max %.2e p99 %.2e iters %.2f gt24 %d" % (d["default"]["max"], d["default"]["p99"], d["default"]["iters_mean"], d["default"]["iters_gt24"]))
PY
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
os.environ["SB_EIG_FP32"] = "1"
print(thth.eta_sweep(cs, tau, fd, etas, edges))
PY
timeout 900 compute-sanitizer --tool racecheck --print-limit 3 python /tmp/race.py 2>&1 | tail -40 > gpurun_out/r2_racecheck_fp32_detail.txt; head -30 gpurun_out/r2_racecheck_fp32_detail.txt | cut -c1-300
for k in 1 2; do timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2c11_ref_$k.json; python - $k <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2c11_ref_%s.json"%sys.argv[1]).read())
print("reference arm run", sys.argv[1], "value %.1f eta-trials/s" % d["value"], "cores", d["cpu_baseline"]["cores"], "single", d["cpu_baseline"]["single_process"]["value"], d["cpu_baseline"]["sample"][:160])
PY
done
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build|row_fft_r2c|tile_fft_tma" -s 12 -c 5 \
    -o gpurun_out/r2c11_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c11_ncu.log 2>&1
