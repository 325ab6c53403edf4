mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_decode.py tests/test_gpu_encode.py tests/test_gpu_reference_api.py -q -m gpu -x -k "bytearray or ByteArray or crc_every or bytes" > gpurun_out/x_tests.txt 2>&1
tail -5 gpurun_out/x_tests.txt
TFR_TRACE=1 timeout 300 python tools/e2e_probe.py 1024 24 both > gpurun_out/x_e2e.txt 2> gpurun_out/x_e2e_trace.txt
grep -v tfr_trace gpurun_out/x_e2e.txt | tail -4
timeout 600 python bench.py --pool 2 --batches-per-step 8 --steps 3 --no-e2e --no-cpu --no-parity --cfg5-passes 0 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/x_bench.json').read().strip().split('\n')[-1])
print(d['value'], {k:(round(v['value'],1), round(v['ms_per_batch'],3)) for k,v in d['extra'].items()})
PY
