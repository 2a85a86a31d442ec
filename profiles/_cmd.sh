timeout 600 python -m pytest tests/test_fused_attention.py tests/test_config_parity.py -m gpu -q 2>&1 | grep -E "d[QKV]|passed|failed" | sort -r | head -14
timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "micro-batch" 
