#!/bin/bash
# round 2, call 15: check warp (9th warp) + balanced row-block shares + hoisted operand loads in
# the tensor-core mat-vec; build kernel: per-CTA fp16 scales, hoisted store offsets; block-level
# reductions (acf_mid / dyn_stats); SB_ROW_DIV=16 A/B for the row FFTs; refreshed ncu captures
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read())
    x=d.get('extra') or {}
    print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4), {k:(round(v.get('device_ms',v.get('per_freq_ms',0)),3), round(v.get('frac',0),4)) for k,v in x.items()})
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
B="--steps 5 --warmup 3 --no-cpu --no-strong"
timeout 400 python bench.py $B 2>gpurun_out/r2c15_a.err | tail -1 > gpurun_out/r2c15_bench_default.json
show "bench default" gpurun_out/r2c15_bench_default.json; tail -2 gpurun_out/r2c15_a.err
SB_ROW_DIV=16 timeout 400 python bench.py $B 2>/dev/null | tail -1 > gpurun_out/r2c15_bench_rowdiv16.json
show "bench ROW_DIV=16" gpurun_out/r2c15_bench_rowdiv16.json
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "passed|failed|error|C3 full|Error|assert|FAILED" | tail -8 > gpurun_out/r2c15_tests.txt
cat gpurun_out/r2c15_tests.txt
timeout 300 python profiles/probe_eig_error.py 2>&1 | tail -1 > gpurun_out/r2c15_eig_error.json; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2c15_eig_error.json"))
    for k in ("default","no_tc"):
        print(k, "vs fp32: max %.2e p99 %.2e iters %.2f" % (d[k]["max"], d[k]["p99"], d[k]["iters_mean"]))
except Exception as ex:
    print("eig_error FAILED", ex)
PY
ncu --set full --clock-control none --import-source on -k regex:"thth_eig_half|thth_build|row_fft_r2c|tile_fft_tma" -s 10 -c 5 \
    -o gpurun_out/r2c15_prof python bench.py --steps 1 --warmup 2 --no-cpu --no-strong --no-extra > gpurun_out/r2c15_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c15_c2_launches.csv python profiles/c2_probe.py > /dev/null 2>&1
SB_ROW_DIV=16 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2c15_c2_launches_rowdiv16.csv python profiles/c2_probe.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/r2c15_c2_launches.csv", "gpurun_out/r2c15_c2_launches_rowdiv16.csv"):
    try:
        rows=[r for r in csv.reader(open(f)) if len(r)>14 and r[0].isdigit()]
        agg=collections.OrderedDict()
        for r in rows:
            agg.setdefault((r[0], r[4][:60]), {})[r[12]]=float(r[14])
        print(f)
        for (i,n),m in list(agg.items())[-9:]:
            if m.get("gpu__time_duration.sum",0) > 8000: print(" ", i, n, "us %.0f" % (m["gpu__time_duration.sum"]/1e3))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
