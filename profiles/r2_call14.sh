#!/bin/bash
# round 2, call 14: tensor-core mat-vec v2 (row-block-major stream of 1 KB units, cheap
# cursor) + shuffle-packed block stores in thth_build_kernel<2>; occupancy-1 experiment
# (SB_EIG_SMEM_PAD: one matrix per SM, 77 MB of fp16 triangles resident in L2)
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
B="--steps 5 --warmup 3 --no-cpu --no-strong --no-extra"
timeout 300 python bench.py $B 2>gpurun_out/r2c14_tc.err | tail -1 > gpurun_out/r2c14_bench_tc.json
show "bench TC v2" gpurun_out/r2c14_bench_tc.json; tail -2 gpurun_out/r2c14_tc.err
SB_EIG_SMEM_PAD=16384 timeout 300 python bench.py $B 2>/dev/null | tail -1 > gpurun_out/r2c14_bench_tc_occ1.json
show "bench TC v2 occ1" gpurun_out/r2c14_bench_tc_occ1.json
SB_EIG_NO_TC=1 SB_EIG_SMEM_PAD=24576 timeout 300 python bench.py $B 2>/dev/null | tail -1 > gpurun_out/r2c14_bench_notc_occ1.json
show "bench NO_TC occ1" gpurun_out/r2c14_bench_notc_occ1.json
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c14_eig_error.json; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c14_eig_error.json"))
    print("fp32 iters", d["fp32_iters_mean"])
    for k in ("default","no_tc"):
        print(k, "vs fp32: max %.2e p99 %.2e iters %.2f gt24 %d" % (d[k]["max"], d[k]["p99"], d[k]["iters_mean"], d[k]["iters_gt24"]), d[k]["worst"][:2])
except Exception as ex:
    print("eig_error FAILED", ex)
PY
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert|FAILED" | tail -8 > gpurun_out/r2c14_tests.txt
cat gpurun_out/r2c14_tests.txt
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build" -s 4 -c 2 \
    -o gpurun_out/r2c14_prof python bench.py --steps 1 --warmup 1 --no-cpu --no-strong --no-extra > gpurun_out/r2c14_ncu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 2>gpurun_out/r2c14_bench.err | tail -1 > gpurun_out/r2c14_bench.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c14_bench.json").read())
    x=d.get('extra') or {}; c=d.get('cpu_baseline') or {}
    print("bench full", round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, 'e2e', round(d['e2e']['value']), 'e2e_f64', d['e2e_f64'] and round(d['e2e_f64']['value']), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3), round(v.get('frac',0),4)) for k,v in x.items()}, 'cpu', c.get('value'), 'err', c.get('max_rel_err_vs_gpu'), 'strong', d.get('strong') and round(d['strong']['value']), 'frac', round(d['roofline']['frac'],4))
except Exception as ex:
    print("bench full FAILED", ex)
PY
