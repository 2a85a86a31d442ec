#!/bin/bash
# GPU check of the fp16-split GEMM: probe (correctness + timing), where-the-time-goes diagnostic, its tests, then the config-5
# bench with AUTO -> F16X3 by the model (1) / never (0)
O=gpurun_out; mkdir -p $O
timeout 420 python profiles/gemm_h16_probe.py ${1:-} > $O/h16_probe.log 2>&1; echo "probe rc=$?" | tee $O/h16_summary.txt
grep -v "^\[h16\]" $O/h16_probe.log | tail -45
timeout 300 python profiles/gemm_h16_prof.py > $O/h16_prof.log 2>&1; echo "prof rc=$?" | tee -a $O/h16_summary.txt
grep -v "^\[h16\]" $O/h16_prof.log; grep "^\[h16\]" $O/h16_prof.log | sort | uniq -c | sort -rn | sed -n '1,2p;9,10p'
timeout 600 python -m pytest tests/test_kernels.py -m gpu -q -x -k "f16x3 or shared_a or guards" > $O/h16_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/h16_summary.txt
tail -3 $O/h16_tests.log
for mode in 1 0; do
B200AD_GEMM_F16X3_AUTO=$mode timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-workloads --no-e2e > $O/h16_bench_$mode.json 2> $O/h16_bench_$mode.err; echo "bench mode $mode rc=$?" | tee -a $O/h16_summary.txt
done
python - <<'PY'
import json
for t in ("1", "0"):
    try:
        l = json.loads(open("gpurun_out/h16_bench_%s.json" % t).read().strip().splitlines()[-1])
        print(t, l["value"], l["ms_per_step"], l["roofline"]["shares"], l.get("clocks"))
    except Exception as e:
        print(t, "no line", e)
PY
