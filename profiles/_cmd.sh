timeout 150 python -m pytest tests/test_rnn_seq.py -m gpu -q -x 2>&1 | tail -12
timeout 120 python profiles/rnn_seq_probe.py 2>&1 | tee gpurun_out/r2d_probe.log
B200AD_RNN_SEQ_PROF=1 timeout 120 python profiles/rnn_seq_probe.py 2>&1 | grep -A10 "rnn_seq prof" | head -24 | tee gpurun_out/r2d_prof.log
