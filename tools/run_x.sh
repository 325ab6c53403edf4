mkdir -p gpurun_out
timeout 300 python tools/quick_encode.py 256 20
timeout 300 python tools/quick_encode.py 1024 10
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/x_enc_launches.csv python tools/quick_encode.py 256 8 > /dev/null 2>&1
python - <<'PY'
import csv
for r in csv.reader(open('gpurun_out/x_enc_launches.csv')):
    if len(r)>10 and r[0].isdigit(): print(r[0], r[4].split('(')[0][:50], r[8], r[7], r[-1])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode_tile_kernel -s 4 -c 1 -f -o gpurun_out/r2_x_encode_tile python tools/quick_encode.py 256 4 > gpurun_out/x_enc_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encode_tile_size -s 4 -c 1 -f -o gpurun_out/r2_x_encode_size python tools/quick_encode.py 256 4 > gpurun_out/x_enc_ncu2.log 2>&1
