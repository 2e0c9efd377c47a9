SB_EIG_PERSIST=96 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sweep_golden or failure or batched or tutorial" 2>&1 | tail -2
for p in 0 64 80 96 112 128 148; do
  SB_EIG_PERSIST=$p timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('persist=$p', round(d['value']), d['ms_per_step'], d['roofline']['kernel_ms']['thth_eig'], d['sweep']['iters_mean'])"
done
