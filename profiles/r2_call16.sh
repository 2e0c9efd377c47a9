#!/bin/bash
# round 2, call 16: software-pipelined tensor-core mat-vec (ldmatrix of unit u+1 behind the MMAs
# of unit u), row-load hoists (constants once per thread, zero padding not loaded, 8-byte
# loads), N/16 row threads as default; racecheck / memcheck of the new kernels
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    x=d.get('extra') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], 'e2e', round(d['e2e']['value']), 'e2e_f64', d.get('e2e_f64') and round(d['e2e_f64']['value']), 'frac', round(d['roofline']['frac'],4), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3), round(v.get('frac',0),4)) for k,v in x.items()})
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong 2>gpurun_out/r2c16_a.err | tail -1 > gpurun_out/r2c16_bench.json
show "bench" gpurun_out/r2c16_bench.json; tail -2 gpurun_out/r2c16_a.err
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert|FAILED" | tail -8 > gpurun_out/r2c16_tests.txt
cat gpurun_out/r2c16_tests.txt
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c16_eig_error.json; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c16_eig_error.json"))
    for k in ("default","no_tc"):
        print(k, "vs fp32: max %.2e p99 %.2e iters %.2f" % (d[k]["max"], d[k]["p99"], d[k]["iters_mean"]))
except Exception as ex:
    print("eig_error FAILED", ex)
PY
timeout 420 compute-sanitizer --tool racecheck --print-limit 3 python profiles/race_sweep.py 2>&1 | tail -12 > gpurun_out/r2_sanitizer_racecheck.txt; cut -c1-200 gpurun_out/r2_sanitizer_racecheck.txt | tail -6
timeout 240 compute-sanitizer --tool memcheck --print-limit 3 python profiles/race_sweep.py 2>&1 | tail -8 > gpurun_out/r2_sanitizer_memcheck_tc.txt; cut -c1-200 gpurun_out/r2_sanitizer_memcheck_tc.txt | tail -3
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build|row_fft_r2c|tile_fft_tma" -s 10 -c 5 \
    -o gpurun_out/r2c16_prof python bench.py --steps 1 --warmup 2 --no-cpu --no-strong --no-extra > gpurun_out/r2c16_ncu.log 2>&1
