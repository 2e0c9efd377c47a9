#!/bin/bash
# round 2, call 19: the shipped code (experimental mat-vec variants removed): bench line,
# GPU tests, racecheck / memcheck, ncu capture, smoke()
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=4 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert|FAILED|s call" | tail -8 > gpurun_out/r2c19_tests.txt
cat gpurun_out/r2c19_tests.txt
timeout 300 compute-sanitizer --tool racecheck --print-limit 3 python profiles/race_sweep.py 2>&1 | tail -8 > gpurun_out/r2_sanitizer_racecheck.txt; cut -c1-200 gpurun_out/r2_sanitizer_racecheck.txt | tail -5
timeout 200 compute-sanitizer --tool memcheck --print-limit 3 python profiles/race_sweep.py 2>&1 | tail -6 > gpurun_out/r2_sanitizer_memcheck_tc.txt; cut -c1-200 gpurun_out/r2_sanitizer_memcheck_tc.txt | tail -2
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build|row_fft_r2c|tile_fft_tma" -s 10 -c 5 \
    -o gpurun_out/r2c19_prof python bench.py --steps 1 --warmup 2 --no-cpu --no-strong --no-extra > gpurun_out/r2c19_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c19_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong --no-extra > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/r2c19_bench.err | tail -1 > gpurun_out/r2c19_bench.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c19_bench.json").read())
    x=d.get('extra') or {}; c=d.get('cpu_baseline') or {}
    print("bench full", round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, 'e2e', round(d['e2e']['value']), 'e2e_f64', d['e2e_f64'] and round(d['e2e_f64']['value']), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3), round(v.get('frac',0),4)) for k,v in x.items()}, 'cpu', c.get('value'), 'err', c.get('max_rel_err_vs_gpu'), 'strong', d.get('strong') and round(d['strong']['value']), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'])
except Exception as ex:
    print("bench full FAILED", ex)
PY
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
