#!/bin/bash
# round 2, call 9: after the lanczos_check fix -- full GPU suite, error probe, bench, ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert" | tail -8 > gpurun_out/r2c9_tests.txt
cat gpurun_out/r2c9_tests.txt
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c9_eig_error.json; python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c9_eig_error.json"))
print("fp32 iters", d["fp32_iters_mean"])
for k,v in d.items():
    if isinstance(v, dict): print(k, "max %.2e p99 %.2e iters %.2f gt24 %d" % (v["max"], v["p99"], v["iters_mean"], v["iters_gt24"]), v["worst"][:2])
PY
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 $EXTRA 2>gpurun_out/r2c9_bench_$label.err | tail -1 > gpurun_out/r2c9_bench_$label.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c9_bench_%s.json"%sys.argv[1]).read())
    x=d.get('extra') or {}
    c=d.get('cpu_baseline') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], d['sweep']['iters_hist'], 'e2e', round(d['e2e']['value']), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3)) for k,v in x.items()}, 'err', c.get('max_rel_err_vs_gpu'), 'strong', d.get('strong') and round(d['strong']['value']))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
EXTRA="" run default SB_X=1
EXTRA="--no-extra --no-cpu --no-strong" run fp32 SB_EIG_FP32=1
EXTRA="--no-extra --no-cpu --no-strong" run etol1e6 SB_EIG_ETOL_B=1e-6
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build|row_fft_r2c|tile_fft_tma|cs_absmax" -s 12 -c 6 \
    -o gpurun_out/r2c9_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c9_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2c9_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong --no-extra > /dev/null 2>&1
