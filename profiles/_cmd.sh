for at in 1 0; do for pb in 0 7; do
  echo "== ATMEM=$at PROBE=$pb"
  env B200AD_TC_ATMEM=$at B200AD_TC_PROBE=$pb timeout 120 python profiles/lstm_step_gemm_probe.py 2>&1 | grep -E "^fwd  h @|^bwd  dA @"
done; done
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max --clock-control none --cache-control none -k regex:gemm_tc_kernel -s 20 -c 4 --csv --log-file gpurun_out/lstm_gemm_ncu.csv python profiles/lstm_step_gemm_probe.py > /dev/null 2>&1
grep -E "^\"" gpurun_out/lstm_gemm_ncu.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | tail -12
