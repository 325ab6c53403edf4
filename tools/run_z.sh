# the round's final measurement on one B200 (outputs under gpurun_out/, copied to profiles/ afterwards)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_z_gputests.txt 2>&1; tail -2 gpurun_out/r2_z_gputests.txt
timeout 900 python bench.py > gpurun_out/r2_z_bench.json 2> gpurun_out/r2_z_bench.err; tail -c 300 gpurun_out/r2_z_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_z_bench_reference_arm.json 2>> gpurun_out/r2_z_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/r2_z_launches.csv python bench.py --steps 1 --warmup 3 --pool 2 --batches-per-step 2 --no-e2e --no-cpu --no-extra --no-parity --cfg5-passes 0 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:decode_tile -s 8 -c 1 -f -o gpurun_out/r2_z_tile_full python tools/quick_resident.py 1024 4 cfg2 > gpurun_out/r2_z_tile_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_bytes -s 8 -c 1 -f -o gpurun_out/r2_z_bytes_tile python tools/quick_resident.py 1024 4 bytes > gpurun_out/r2_z_bytes_ncu.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py tests/test_gpu_reference_api.py tests/test_emulator.py -q -m gpu -x -k "bytearray or consumed or emulator or streaming or read" > gpurun_out/r2_z_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_z_memcheck.log; tail -4 gpurun_out/r2_z_memcheck.log
