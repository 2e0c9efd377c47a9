#!/bin/bash
# round 2, call 6: batched-gather build kernel, leaner cp.async fetch, ACF TMA, extended C3 check
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_random.py -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full" | tail -8 > gpurun_out/r2c6_tests.txt
cat gpurun_out/r2c6_tests.txt
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-strong $EXTRA 2>gpurun_out/r2c6_bench_$label.err | tail -1 > gpurun_out/r2c6_bench_$label.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c6_bench_%s.json"%sys.argv[1]).read())
    x=d.get('extra') or {}
    c=d.get('cpu_baseline') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], d['sweep']['iters_hist'], 'e2e', round(d['e2e']['value']), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3)) for k,v in x.items()}, 'err', c.get('max_rel_err_vs_gpu'))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
EXTRA="" run default SB_X=1
EXTRA="--no-extra" run etol1e6 SB_EIG_ETOL_B=1e-6
EXTRA="--no-extra --no-cpu" run rtol1e3 SB_EIG_RTOL_R=1e-3
EXTRA="--no-extra --no-cpu" run rtol5e3 SB_EIG_RTOL_R=5e-3
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_bf16|thth_build" -s 6 -c 2 \
    -o gpurun_out/r2c6_sweep python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c6_ncu.log 2>&1
