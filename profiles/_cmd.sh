B200AD_PARITY_REPORT=gpurun_out/r2e_parity_ledger.json timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2e_pytest_gpu.log | tail
timeout 300 python bench.py --workload c4 > gpurun_out/r2e_c4.json 2> gpurun_out/r2e_c4.err; echo "c4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2e_c4.json').read().strip().splitlines()[-1]); print(json.dumps(d['cells'],indent=0)); print(d.get('clocks'))"
timeout 300 python bench.py --workload c1 > gpurun_out/r2e_c1.json 2> gpurun_out/r2e_c1.err; echo "c1 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2e_c1.json').read().strip().splitlines()[-1]); print(json.dumps(d['through_api'],indent=0))"
