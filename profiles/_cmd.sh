L=auto_differentiation_b200/libb200ad.so
cp $L /tmp/orig.so
for rep in 1 2; do
for v in nohint hint; do
  cp profiles/_lib_$v.so $L
  echo "== $v"; timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "micro-batch"
  timeout 200 python profiles/gemm_tc_probe.py 2>&1 | grep -iE "tflop|ms" | head -6
done
done
cp profiles/_lib_hint.so $L
timeout 300 python -m pytest tests/test_rnn_seq.py tests/test_fused_attention.py -m gpu -q 2>&1 | tail -1
cp /tmp/orig.so $L
