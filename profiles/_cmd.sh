timeout 600 python profiles/attn_stress.py 2>&1 | tail -8
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -2; done
