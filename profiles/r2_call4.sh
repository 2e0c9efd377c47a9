#!/bin/bash
# round 2, call 4: eig_bf16 v2 (scalar-q FFMA2, deferred check), TMA tile loads + column chunks in the FFT passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_random.py tests/test_gpu_retrieval.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2c4_tests.txt
cat gpurun_out/r2c4_tests.txt
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong 2>gpurun_out/r2c4_bench_$label.err | tail -1 > gpurun_out/r2c4_bench_$label.json
  python - "$label" <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/r2c4_bench_%s.json"%sys.argv[1]).read())
    x=d.get('extra') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], 'e2e', round(d['e2e']['value']), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3)) for k,v in x.items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run default SB_X=1
run notma SB_FFT_NO_TMA=1
run nochunk SB_COL_CHUNK_MB=0
run chunk24 SB_COL_CHUNK_MB=24
run etol1e6 SB_EIG_ETOL_B=1e-6
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_bf16|tile_fft_tma" -s 8 -c 6 \
    -o gpurun_out/r2c4_prof python bench.py --steps 1 --warmup 3 --no-cpu --no-strong --no-extra > gpurun_out/r2c4_ncu.log 2>&1
