#!/bin/bash
# round 2 multi-GPU call: gpurun --gpus N -- 'bash profiles/r2_multigpu.sh N'
# bench (weak + strong legs) and BASELINE configs 4 / 5 on N GPUs of one node.
N=${1:-8}
mkdir -p gpurun_out
PORT=29517
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extra > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/r2_bench_%sgpu.json"%n).read().strip().splitlines()[-1])
    print("bench", n, "gpus: weak", round(d['value']), round(d['ms_per_step'],3), "strong", round(d['strong']['value']), round(d['strong']['ms_per_step'],3), "e2e", round(d['e2e']['value']), d['roofline']['kernel_ms'])
except Exception as e:
    print("bench FAILED", e); print(open("gpurun_out/r2_bench_%sgpu.err"%n).read()[-1500:])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    profiles/multi_gpu_c4_c5.py --nreal 256 --ns 8192 --nf 64 --ndyn 1024 --neta 256 --out gpurun_out/r2_c4c5_${N}gpu.json 2> gpurun_out/r2_c4c5_${N}gpu.err | tail -2
tail -3 gpurun_out/r2_c4c5_${N}gpu.err
