#!/bin/bash
# 8-GPU call, bench only: the scaling lines at N = 8, 4, 2 (each with its dp_parity check)
O=gpurun_out; mkdir -p $O
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --steps 5 --warmup 3 > $O/r2_bench_n$n.json 2> $O/r2_bench_n$n.err; echo "bench$n rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('$O/r2_bench_n$n.json').read().splitlines() if l.startswith('{')][-1])
print($n, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['ms_per_step'], d.get('dp_parity',{}).get('ok'), d.get('dp_parity',{}).get('worst_err_over_tol'))
PY
done
