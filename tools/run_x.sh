mkdir -p gpurun_out
for cw in 8 10 12 16; do echo "cw $cw"; TFR_BYTES_CW=$cw timeout 300 python tools/quick_resident.py 1024 100 bytes 2>&1 | grep submit | cut -c1-200; done
echo "cw 8 nofill"; TFR_BYTES_NOFILL=1 TFR_BYTES_CW=8 timeout 300 python tools/quick_resident.py 1024 100 bytes 2>&1 | grep submit | cut -c1-200
echo "cw 12 nofill"; TFR_BYTES_NOFILL=1 TFR_BYTES_CW=12 timeout 300 python tools/quick_resident.py 1024 100 bytes 2>&1 | grep submit | cut -c1-200
echo "cw 8 noprefetch"; TFR_NO_L2_PREFETCH=1 TFR_BYTES_CW=8 timeout 300 python tools/quick_resident.py 1024 100 bytes 2>&1 | grep submit | cut -c1-200
TFR_BYTES_CW=8 timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_bytes -s 6 -c 1 -f -o gpurun_out/r2_x_bytes_tile python tools/quick_resident.py 1024 4 bytes > gpurun_out/x_bytes_ncu.log 2>&1
