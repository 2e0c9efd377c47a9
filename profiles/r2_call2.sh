#!/bin/bash
# round 2, call 2: new default eigen solver (eig_bf16.cu) -- parity, A/B bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c2_tests.txt
cat gpurun_out/r2c2_tests.txt
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu 2>gpurun_out/r2c2_bench_$label.err | tail -1 > gpurun_out/r2c2_bench_$label.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c2_bench_%s.json"%sys.argv[1]).read())
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep'], round(d['e2e']['value']))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default SB_X=1
run fp32 SB_EIG_FP32=1
run etol1e6 SB_EIG_ETOL_B=1e-6
run etol5e6 SB_EIG_ETOL_B=5e-6
ncu --set full --clock-control none --import-source on -k regex:thth_eig_bf16 -s 3 -c 1 \
    -o gpurun_out/r2c2_eig python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2c2_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r2c2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
