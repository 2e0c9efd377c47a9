#!/bin/bash
# round 2, call 5: eig_bf16 cp.async vs bulk fetch, f2 GPU tests, multi-GPU script sanity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_arcfit.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2c5_tests.txt
cat gpurun_out/r2c5_tests.txt
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-extra 2>gpurun_out/r2c5_bench_$label.err | tail -1 > gpurun_out/r2c5_bench_$label.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c5_bench_%s.json"%sys.argv[1]).read())
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], d['sweep']['iters_max'], 'e2e', round(d['e2e']['value']))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run cpasync SB_X=1
run bulk SB_EIG_BULK=1
run cpasync_etol1e6 SB_EIG_ETOL_B=1e-6
timeout 300 python profiles/multi_gpu_c4_c5.py --nreal 4 --ns 1024 --nf 8 --ndyn 4 --neta 64 --out gpurun_out/r2c5_c4c5_tiny.json 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_bf16" -s 3 -c 1 \
    -o gpurun_out/r2c5_eig_cpa python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c5_ncu.log 2>&1
SB_EIG_BULK=1 ncu --set full --clock-control none --import-source on -k regex:"thth_eig_bf16" -s 3 -c 1 \
    -o gpurun_out/r2c5_eig_bulk python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c5_ncu2.log 2>&1
