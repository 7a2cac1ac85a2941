for mode in 0 1 0 1; do
  ALIGNSDF_BENCH_STEP_SYNC=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 12 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step_sync=$mode', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
