for mode in clocks noclocks clocks noclocks; do
  if [ $mode = noclocks ]; then export ALIGNSDF_BENCH_NO_CLOCKS=1; else unset ALIGNSDF_BENCH_NO_CLOCKS; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 16 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'])"
done
