B200AD_ATTN_PROF=1 timeout 120 python profiles/attn_probe.py timing 2>&1 | grep ah_prof | tail -3
