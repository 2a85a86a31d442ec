#!/bin/bash
# One gpurun call: topology, the -m gpu suite (with the parity ledger), the default bench line, the memory-bound ncu pass
# and the single-GPU H2D probe.  Outputs land in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash profiles/run_gpu_checks.sh r2a'
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
{ nvidia-smi topo -m; lscpu | head -30; for f in /sys/devices/system/node/node*/cpulist; do echo "$f: $(cat $f)"; done; free -g; nproc; } > $O/${TAG}_topology.txt 2>&1
B200AD_PARITY_REPORT=$O/${TAG}_parity_ledger.json timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/${TAG}_pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench rc=$?" >> $O/${TAG}_bench.err
timeout 200 python profiles/membound_kernels.py > $O/${TAG}_membound_events.json 2> $O/${TAG}_membound.err &&
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'unary_|binary_|softmax_|xent_|adam_multi|sgd_multi|colsum|rowsum|reduce_narrow|lstm_cell|embedding_|fill_kernel' \
    --csv --log-file $O/${TAG}_membound_ncu.csv python profiles/membound_kernels.py > $O/${TAG}_membound_ncu.log 2>&1
timeout 200 python profiles/h2d_probe.py > $O/${TAG}_h2d_probe_n1.json 2> $O/${TAG}_h2d_probe.err
tail -5 $O/${TAG}_pytest_gpu.log; tail -c 600 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
