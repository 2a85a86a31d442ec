#!/bin/bash
# 8-GPU call: bare H2D probe (alone / together, allocation variants) + the scaling bench at N=8 (and N=4, N=2 if time allows)
O=gpurun_out; mkdir -p $O
{ nvidia-smi topo -m; lscpu | grep -E "^CPU\(s\)|NUMA|Socket|Model name"; for f in /sys/devices/system/node/node*/cpulist; do echo "$f: $(cat $f)"; done; free -g | head -2; } > $O/r2_topology_n8.txt 2>&1
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 profiles/h2d_probe.py > $O/r2_h2d_probe_n8.json 2> $O/r2_h2d_probe_n8.err; echo "probe rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err; echo "bench8 rc=$?"
B200AD_H2D_STREAMS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-dp-parity > $O/r2_bench_n8_s1.json 2> $O/r2_bench_n8_s1.err; echo "bench8 1-stream rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/dp_parity_nccl.py > $O/r2_dp_parity_n2.log 2>&1; echo "dp parity rc=$?"
tail -c 1200 $O/r2_bench_n8.json; tail -3 $O/r2_bench_n8.err; tail -c 600 $O/r2_h2d_probe_n8.json
