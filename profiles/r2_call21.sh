#!/bin/bash
# round 2, call 21 (last GPU minutes): fused Lanczos step of the tensor-core path (4 barriers per
# step instead of 7): quick bench line, GPU tests, racecheck
mkdir -p gpurun_out
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu --no-strong --no-extra 2>gpurun_out/r2c21.err | tail -1 > gpurun_out/r2c21_bench.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2c21_bench.json").read())
    print("bench", round(d['value']), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()}, d['sweep']['iters_mean'], 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],4))
except Exception as ex:
    print("bench FAILED", ex); print(open("gpurun_out/r2c21.err").read()[-800:])
PY
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|error|Error|assert|FAILED" | tail -5 > gpurun_out/r2c21_tests.txt
cat gpurun_out/r2c21_tests.txt
timeout 60 compute-sanitizer --tool racecheck --print-limit 3 python profiles/race_sweep.py 2>&1 | tail -6 > gpurun_out/r2c21_racecheck.txt; cut -c1-200 gpurun_out/r2c21_racecheck.txt | tail -4
