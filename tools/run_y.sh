mkdir -p gpurun_out
python bench.py --pool 4 --batches-per-step 16 --steps 4 --warmup 3 --no-cpu --no-extra --no-parity --cfg5-passes 0 > gpurun_out/r2_y_bench_e2e.json 2> gpurun_out/r2_y_bench_e2e.err
timeout 480 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decode.py tests/test_gpu_scale.py -q -m gpu -x -k "not steady_state and not pipelined and not 250k and not 100k and not cfg2" > gpurun_out/r2_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py tests/test_gpu_decode.py -q -m gpu -x -k "ragged_one_pass or malformed or speculation or roundtrip" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_racecheck.log
for k in ragged seq strings; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_tile -s 6 -c 1 -f -o gpurun_out/r2_y_${k}_tile python tools/quick_resident.py 256 4 $k > gpurun_out/r2_y_${k}_ncu.log 2>&1
done
tail -3 gpurun_out/r2_memcheck.log gpurun_out/r2_racecheck.log; cat gpurun_out/r2_y_bench_e2e.json | head -c 3000
