for pb in 0 1 2 3 4 7; do
  echo "== PROBE=$pb"
  env B200AD_TC_PROBE=$pb timeout 120 python profiles/lstm_step_gemm_probe.py 2>&1 | grep -E "^fwd  h @|^bwd  dA @"
done
