timeout 600 python -m pytest tests/test_fused_attention.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error|d_k=" | sort -r | head -12
