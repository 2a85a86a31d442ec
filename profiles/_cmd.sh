timeout 600 python -m pytest tests/test_rnn_seq.py tests/test_config_parity.py tests/test_models_and_dp.py tests/test_layers_parity.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -4
for v in 0 1; do
echo "== NO_PDL=$v"
if [ $v = 1 ]; then export B200AD_LSTM_NO_PDL=1; fi
timeout 300 python bench.py --workload c4 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps(d.get('cells')['lstm']))"
done
