timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -5
for w in c4 c3 c1; do timeout 300 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/r2m_$w.json; done
python - <<'PY'
import json
for w in ("c4","c3","c1"):
    d=json.loads(open('gpurun_out/r2m_%s.json'%w).read().strip().splitlines()[-1])
    print(w, d.get('value'), json.dumps(d.get("through_api"))[:900])
PY
