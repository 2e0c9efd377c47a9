#!/bin/bash
# round 2, call 20 (gpurun --gpus 2): the shipped code under torchrun on two GPUs -- weak and
# strong legs of bench.py (one NCCL all-gather of eigenvalues per step)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r2_bench_2gpu_final.json 2> gpurun_out/r2_bench_2gpu_final.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_2gpu_final.json").read().strip().splitlines()[-1])
    print("bench 2 gpus: weak", round(d['value']), round(d['ms_per_step'],3), "strong", round(d['strong']['value']), round(d['strong']['ms_per_step'],3), "e2e", round(d['e2e']['value']), {k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
except Exception as e:
    print("bench FAILED", e); print(open("gpurun_out/r2_bench_2gpu_final.err").read()[-1500:])
PY
