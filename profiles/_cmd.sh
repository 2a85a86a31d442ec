timeout 150 python -m pytest tests/test_rnn_seq.py -m gpu -q -x 2>&1 | tail -4
timeout 120 python profiles/rnn_seq_probe.py 2>&1 | tail -2
B200AD_RNN_SEQ_PROF=1 timeout 120 python profiles/rnn_seq_probe.py 2>&1 | grep -A10 "rnn_seq prof" | head -11 | tee gpurun_out/r2h_prof.log
