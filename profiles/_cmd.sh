echo "== timing h16 with prof"; B200AD_ATTN_PROF=1 timeout 120 python profiles/attn_probe.py timing 2>&1 | grep -E "ah_prof|config-5" | tail -14
