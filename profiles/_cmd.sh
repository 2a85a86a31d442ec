L=auto_differentiation_b200/libb200ad.so
cp $L /tmp/orig.so
for rep in 1 2; do
for v in head regs160 defer; do
  cp profiles/_lib_$v.so $L
  echo -n "$v: "; timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "micro-batch bwd"
done
done
cp profiles/_lib_defer.so $L
timeout 300 python -m pytest tests/test_fused_attention.py -m gpu -q 2>&1 | tail -1
cp /tmp/orig.so $L
