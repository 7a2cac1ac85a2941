for rep in 1 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 16 --warmup 3 2>/dev/null | tee gpurun_out/bench_4gpu_c$rep.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rep $rep', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['clocks'])"
done
