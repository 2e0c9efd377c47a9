#!/bin/bash
# round 2, call 18: per-lane cp.async ring (EB_MODE_TC) against cp.async.bulk ring (EB_MODE_TCB), both on the
# pre-swizzled block layout; the faster one is then validated: error probe over all 1024 curvatures,
# GPU tests, racecheck / memcheck, ncu capture, the full bench line, smoke()
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    x=d.get('extra') or {}; c=d.get('cpu_baseline') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], 'e2e', round(d['e2e']['value']), 'e2e_f64', d.get('e2e_f64') and round(d['e2e_f64']['value']), 'frac', round(d['roofline']['frac'],4), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3), round(v.get('frac',0),4)) for k,v in x.items()}, 'cpu', c.get('value'), 'err', c.get('max_rel_err_vs_gpu'), 'strong', d.get('strong') and round(d['strong']['value']))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
B="--steps 5 --warmup 3 --no-cpu --no-strong --no-extra"
SB_EIG_TCBULK=0 timeout 300 python bench.py $B 2>gpurun_out/r2c18_a.err | tail -1 > gpurun_out/r2c18_bench_cpa.json
show "bench cp.async" gpurun_out/r2c18_bench_cpa.json; tail -2 gpurun_out/r2c18_a.err
SB_EIG_TCBULK=1 timeout 300 python bench.py $B 2>/dev/null | tail -1 > gpurun_out/r2c18_bench_bulk.json
show "bench bulk" gpurun_out/r2c18_bench_bulk.json
CH=$(python - <<'PY'
import json
try:
    a=json.loads(open("gpurun_out/r2c18_bench_cpa.json").read())["roofline"]["kernel_ms"]["thth_eig"]
    b=json.loads(open("gpurun_out/r2c18_bench_bulk.json").read())["roofline"]["kernel_ms"]["thth_eig"]
    print(1 if b < 0.985 * a else 0)
except Exception:
    print(0)
PY
)
echo "chosen SB_EIG_TCBULK=$CH" | tee gpurun_out/r2c18_choice.txt
export SB_EIG_TCBULK=$CH
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c18_eig_error.json; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c18_eig_error.json"))
    for k in ("default","no_tc","pair","tcbulk","tccpa"):
        print(k, "vs fp32: max %.2e p99 %.2e iters %.2f" % (d[k]["max"], d[k]["p99"], d[k]["iters_mean"]))
except Exception as ex:
    print("eig_error FAILED", ex)
PY
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=6 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert|FAILED|s call|s setup" | tail -14 > gpurun_out/r2c18_tests.txt
cat gpurun_out/r2c18_tests.txt
timeout 300 compute-sanitizer --tool racecheck --print-limit 3 python profiles/race_sweep.py 2>&1 | tail -12 > gpurun_out/r2_sanitizer_racecheck.txt; cut -c1-200 gpurun_out/r2_sanitizer_racecheck.txt | tail -3
timeout 200 compute-sanitizer --tool memcheck --print-limit 3 python profiles/race_sweep.py 2>&1 | tail -8 > gpurun_out/r2_sanitizer_memcheck_tc.txt; cut -c1-200 gpurun_out/r2_sanitizer_memcheck_tc.txt | tail -2
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build|row_fft_r2c|tile_fft_tma" -s 10 -c 5 \
    -o gpurun_out/r2c18_prof python bench.py --steps 1 --warmup 2 --no-cpu --no-strong --no-extra > gpurun_out/r2c18_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c18_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong --no-extra > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/r2c18_bench.err | tail -1 > gpurun_out/r2c18_bench.json
show "bench full" gpurun_out/r2c18_bench.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
