timeout 600 python -m pytest tests/test_fused_attention.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "micro-batch" 
B200AD_ATTN_H16=1 timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "micro-batch" 
B200AD_ATTN_PROF=1 timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "ah_prof" | head -2
