#!/bin/bash
# GPU check of the fp16-split GEMM: probe (correctness + timing), its tests, then the config-5 bench with AUTO -> F16X3 on / off
O=gpurun_out; mkdir -p $O
timeout 420 python profiles/gemm_h16_probe.py ${1:-} > $O/h16_probe.log 2>&1; echo "probe rc=$?" | tee $O/h16_summary.txt
tail -40 $O/h16_probe.log
timeout 600 python -m pytest tests/test_kernels.py -m gpu -q -x -k "f16x3 or shared_a or guards" > $O/h16_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/h16_summary.txt
tail -15 $O/h16_tests.log
B200AD_GEMM_F16X3_AUTO=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-e2e > $O/h16_bench_on.json 2> $O/h16_bench_on.err; echo "bench on rc=$?" | tee -a $O/h16_summary.txt
B200AD_GEMM_F16X3_AUTO=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-e2e > $O/h16_bench_off.json 2> $O/h16_bench_off.err; echo "bench off rc=$?" | tee -a $O/h16_summary.txt
python - <<'PY'
import json
for t in ("on", "off"):
    try:
        l = json.loads(open("gpurun_out/h16_bench_%s.json" % t).read().strip().splitlines()[-1])
        print(t, l["value"], l["ms_per_step"], l["roofline"]["shares"], l.get("clocks"))
    except Exception as e:
        print(t, "no line", e)
PY
