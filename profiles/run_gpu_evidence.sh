#!/bin/bash
# One gpurun call: the evidence files for the tensor-core kernels (round 2, after the fused dK+dV attention pass).
#   gpurun --timeout 1800 -- 'bash profiles/run_gpu_evidence.sh'
O=gpurun_out
mkdir -p $O
timeout 60 ./profiles/alu_rate > $O/r2_alu_rate.log 2>&1
timeout 60 ./profiles/mma_rate2 > $O/r2_mma_rate2.log 2>&1
# role cycle counters + one item's event trace of the attention backward passes (256 sequences)
B200AD_ATTN_PROF=2 timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "ah_prof|ah_trace|micro-batch" | awk '!seen[$0]++' | head -120 > $O/r2_attn_h16_trace.log
timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "micro-batch" > $O/r2_attn_timing.log
# per-launch duration / DRAM bytes / issue statistics of the attention passes at the config-5 micro-batch (256 sequences)
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max \
    --clock-control none -k regex:'attn' -s 3 -c 8 --csv --log-file $O/r2_attn_passes_ncu.csv python profiles/attn_probe.py timing > $O/attn_ncu.log 2>&1
# --set full of every tensor-core kernel family (small shapes), exported on the box
python profiles/ncu_targets.py > $O/plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc3|attn_tc_kernel|attn_h16_kernel|attn_amax|rnn_seq_kernel' \
    -o $O/r2_tensor_kernels -f python profiles/ncu_targets.py > $O/ncu_full.log 2>&1
ncu -i $O/r2_tensor_kernels.ncu-rep --page raw --csv > $O/r2_ncu_tensor_kernels_raw.csv 2>/dev/null
python profiles/ncu_select.py $O/r2_ncu_tensor_kernels_raw.csv > $O/r2_ncu_tensor_kernels.csv
rm -f $O/r2_tensor_kernels.ncu-rep
# launch list of the bench command (one warm-up + one timed step)
timeout 300 python bench.py --steps 1 --warmup 1 --no-workloads --no-cpu-baseline > $O/bench_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/r2_launches_c5.csv \
    python bench.py --steps 1 --warmup 1 --no-workloads --no-cpu-baseline > $O/bench_ncu.log 2>&1
ls -la $O | tail -20
cat $O/r2_attn_timing.log
