#!/bin/bash
# Round-2 second-session evidence: ncu --set full of the fp16-split GEMM, launch list of the bench command, final bench line.
O=gpurun_out; mkdir -p $O
python profiles/ncu_targets_h16.py > $O/plain_h16.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_h16_kernel|h16_split|h16_rowscale' \
    -o $O/r2b_h16 -f python profiles/ncu_targets_h16.py > $O/ncu_h16.log 2>&1
ncu -i $O/r2b_h16.ncu-rep --page raw --csv > $O/r2b_ncu_gemm_h16_raw.csv 2>/dev/null
python profiles/ncu_select.py $O/r2b_ncu_gemm_h16_raw.csv > $O/r2b_ncu_gemm_h16.csv
rm -f $O/r2b_h16.ncu-rep
cat $O/r2b_ncu_gemm_h16.csv | cut -c1-400
timeout 200 python bench.py --steps 1 --warmup 1 --no-workloads --no-cpu-baseline > $O/bench_plain.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2b_launches_c5.csv \
    python bench.py --steps 1 --warmup 1 --no-workloads --no-cpu-baseline > $O/bench_ncu.log 2>&1
timeout 150 python profiles/gemm_h16_probe.py > $O/r2b_gemm_h16_probe.log 2>&1; echo "probe rc=$?"
timeout 150 python profiles/gemm_h16_prof.py > $O/r2b_gemm_h16_prof.log 2>&1; echo "prof rc=$?"
B200AD_PARITY_REPORT=$O/r2b_parity_ledger.json timeout 600 python -m pytest tests -m gpu -q > $O/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" $O/r2b_pytest_gpu.log | tail -3
timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2b_bench.json 2> $O/r2b_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16 --no-workloads --no-cpu-baseline > $O/r2b_bench_bf16.json 2> $O/r2b_bench_bf16.err; echo "bf16 bench rc=$?"
python - <<'PY'
import json
for f in ("r2b_bench", "r2b_bench_bf16"):
    l = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, l["value"], l["ms_per_step"], l["e2e"]["value"], l["roofline"]["avg_launch_ms"], l["roofline"]["frac"], l["roofline"]["shares"], l.get("clocks"))
    for k, v in (l.get("workloads") or {}).items():
        print("  ", k, v.get("value"), v.get("unit"), (v.get("through_api") or {}).get("steps_per_s"), (v.get("through_api") or {}).get("graphed_steps_per_s"))
PY
