#!/bin/bash
# Round-2 (second session) validation: probe of the fp16-split GEMM, the whole -m gpu suite with the parity ledger, the default
# bench line (with the workloads block), and the ncu --set full capture of gemm_h16_kernel at the shared-A projection shape and 8192^3.
TAG=${1:-r2b}
O=gpurun_out; mkdir -p $O
timeout 420 python profiles/gemm_h16_probe.py > $O/${TAG}_gemm_h16_probe.log 2>&1; echo "probe rc=$?" | tee $O/${TAG}_summary.txt
grep -E "square|shared|c4|fwd" $O/${TAG}_gemm_h16_probe.log
B200AD_PARITY_REPORT=$O/${TAG}_parity_ledger.json timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_summary.txt
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest_gpu.log | tail -12
timeout 600 python bench.py --steps 5 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?" | tee -a $O/${TAG}_summary.txt
tail -c 3000 $O/${TAG}_bench.json
