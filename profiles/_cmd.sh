timeout 600 python -m pytest tests/test_fused_attention.py tests/test_layers_parity.py tests/test_config_parity.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -3
for i in 1 2; do timeout 300 python bench.py --workload c3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps(d.get('through_api')))"; done
timeout 300 python bench.py --workload c1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps(d.get('through_api'))[:600])"
