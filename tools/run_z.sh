# the round's final measurement on one B200 (outputs under gpurun_out/, copied to profiles/ afterwards)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_z_gputests.txt 2>&1; tail -2 gpurun_out/r2_z_gputests.txt
timeout 900 python bench.py > gpurun_out/r2_z_bench.json 2> gpurun_out/r2_z_bench.err; tail -c 300 gpurun_out/r2_z_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_z_bench_reference_arm.json 2>> gpurun_out/r2_z_bench.err
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_encode.py -q -m gpu -x > gpurun_out/r2_z_memcheck_encode.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_z_memcheck_encode.log; tail -3 gpurun_out/r2_z_memcheck_encode.log
