mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/x_tests.txt 2>&1
tail -4 gpurun_out/x_tests.txt
timeout 900 python bench.py --pool 4 --batches-per-step 16 --steps 4 --no-cpu --no-parity --cfg5-passes 1 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err
tail -3 gpurun_out/x_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/x_bench.json').read().strip().split('\n')[-1])
print(d['value'], d['e2e']['value'], d['e2e'].get('one_handle_one_thread'), d['e2e'].get('two_handles_two_threads'))
print({k:(round(v['value'],1), round(v['ms_per_batch'],3)) for k,v in d['extra'].items()})
print(d['cfg5_file_sharded']['value'], d['cfg5_file_sharded'].get('blocks'))
PY
