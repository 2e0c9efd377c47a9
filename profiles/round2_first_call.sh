#!/bin/bash
# First gpurun call of round 2: validates what round 1 could only compile.
#   gpurun --timeout 1200 -- 'bash profiles/round2_first_call.sh'
mkdir -p gpurun_out
SB_TEST_UNVERIFIED=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_retrieval.py \
    -m gpu -q -k "mixed_precision or search_batch or scale_dyn or non_power_of_two" 2>&1 | tail -25 \
    > gpurun_out/r2_tests.txt
cat gpurun_out/r2_tests.txt
for v in 0 1 2; do
  SB_EIG_MIXED=$v timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('SB_EIG_MIXED=$v', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms'], d['sweep'])"
done | tee gpurun_out/r2_mixed_bench.txt
timeout 300 python profiles/probe_search_batch.py 2>&1 | tail -1 | tee gpurun_out/r2_search_batch.json
