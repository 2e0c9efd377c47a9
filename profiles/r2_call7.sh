#!/bin/bash
# round 2, call 7: fp16 (scaled) iteration copy, packed-fp32 FFT butterflies, error probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert" | tail -8 > gpurun_out/r2c7_tests.txt
cat gpurun_out/r2c7_tests.txt
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c7_eig_error.json; cat gpurun_out/r2c7_eig_error.json | cut -c1-2500
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-strong $EXTRA 2>gpurun_out/r2c7_bench_$label.err | tail -1 > gpurun_out/r2c7_bench_$label.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c7_bench_%s.json"%sys.argv[1]).read())
    x=d.get('extra') or {}
    c=d.get('cpu_baseline') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], d['sweep']['iters_hist'], 'e2e', round(d['e2e']['value']), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3)) for k,v in x.items()}, 'err', c.get('max_rel_err_vs_gpu'))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
EXTRA="" run default SB_X=1
EXTRA="--no-extra --no-cpu" run rows8 SB_BUILD_ROWS=8
EXTRA="--no-extra --no-cpu" run etol1e6 SB_EIG_ETOL_B=1e-6
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|row_fft_r2c|tile_fft_tma|cs_absmax" -s 10 -c 5 \
    -o gpurun_out/r2c7_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c7_ncu.log 2>&1
