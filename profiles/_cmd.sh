timeout 120 python profiles/lstm_step_gemm_probe.py 2>&1 | tail -8
