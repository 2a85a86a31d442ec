timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED" | tail -3
for w in c3 c1 c4; do timeout 300 python bench.py --workload $w > gpurun_out/r2j_$w.json 2>/dev/null; done
python - <<'PY'
import json
for w in ("c3","c1","c4"):
    d=json.loads(open('gpurun_out/r2j_%s.json'%w).read().strip().splitlines()[-1])
    print(w, json.dumps(d.get("through_api"))[:700])
PY
