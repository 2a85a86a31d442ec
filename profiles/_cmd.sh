timeout 600 python -m pytest tests/test_fused_attention.py tests/test_rnn_seq.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -6
