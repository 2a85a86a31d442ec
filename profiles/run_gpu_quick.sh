#!/bin/bash
# quick GPU iteration: one risky test file under its own timeout, then the whole -m gpu suite, then side workloads
TAG=${1:-q}; shift
O=gpurun_out; mkdir -p $O
for f in "$@"; do
  timeout 180 python -m pytest $f -m gpu -q -x > $O/${TAG}_$(basename $f .py).log 2>&1; echo "$f rc=$?" | tee -a $O/${TAG}_summary.txt
done
B200AD_PARITY_REPORT=$O/${TAG}_parity_ledger.json timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest_gpu.log 2>&1; echo "full suite rc=$?" | tee -a $O/${TAG}_summary.txt
timeout 300 python bench.py --workload c4 > $O/${TAG}_c4.json 2> $O/${TAG}_c4.err; echo "c4 rc=$?" | tee -a $O/${TAG}_summary.txt
tail -c 1500 $O/${TAG}_c4.json
grep -E "passed|failed|FAILED|Error" $O/${TAG}_pytest_gpu.log | tail -15
