#!/bin/bash
# round 2, call 3: FFMA2 eigen kernel + hoisted build kernel + pipelined search_batch
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_random.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c3_tests.txt
cat gpurun_out/r2c3_tests.txt
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-extra 2>gpurun_out/r2c3_bench_$label.err | tail -1 > gpurun_out/r2c3_bench_$label.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c3_bench_%s.json"%sys.argv[1]).read())
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep'], 'e2e', round(d['e2e']['value']), 'strong', d['strong'] and round(d['strong']['value']))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default SB_X=1
run fp32 SB_EIG_FP32=1
run etol1e6 SB_EIG_ETOL_B=1e-6
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c3_bench_full.json 2> gpurun_out/r2c3_bench_full.err
tail -c 3000 gpurun_out/r2c3_bench_full.json
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_bf16|thth_build" -s 6 -c 2 \
    -o gpurun_out/r2c3_sweep python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c3_ncu.log 2>&1
