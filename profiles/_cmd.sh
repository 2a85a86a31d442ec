echo "== correctness"; timeout 150 python profiles/attn_probe.py quick 2>&1 | tail -24 | grep -E "err" | awk '{print $NF}' | sort -g | tail -3
echo "== default"; timeout 120 python profiles/attn_probe.py timing 2>&1 | tail -2
echo "== prof"; B200AD_ATTN_PROF=1 timeout 120 python profiles/attn_probe.py timing 2>&1 | grep ah_prof | tail -3
timeout 100 python -m pytest tests/test_fused_attention.py -m gpu -q 2>&1 | tail -2
