#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are made here by
# tools/summarize_profiles.py).  Numbers printed by programs under ncu are never bench values.
set -u
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
# 1. launch list of the bench command itself (headline step, n = 1e9)
timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 10 --warmup 3 --no-side > gpurun_out/launches_r02.out 2>&1
# 2. the same step at BASELINE configs[1] (n = 1e8)
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_r02.csv \
    python bench.py --n 1e8 --steps 10 --warmup 3 --no-side > gpurun_out/launches_cfg2_r02.out 2>&1
# 3. --set full of the dominant kernel at both sizes
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cnsm_stream -s 3 -c 1 -f -o gpurun_out/stream_1e8_r02 \
    python tools/one_query.py 1e8 1024 2048 5 5 > gpurun_out/stream_1e8_r02.out 2>&1
timeout 900 ncu --set full --clock-control none -k regex:cnsm_stream -s 2 -c 1 -f -o gpurun_out/stream_1e9_r02 \
    python tools/one_query.py 1e9 1024 2048 5 4 > gpurun_out/stream_1e9_r02.out 2>&1
# 4. window-mean pass (all kernels of one call) and the cNSM-DTW stages
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wmean|rewalk_batch' -s 2 -c 2 -f -o gpurun_out/wmean_1e8_r02 \
    python tools/one_wmean.py 1e8 2 > gpurun_out/wmean_1e8_r02.out 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/launches_wmean_r02.csv \
    python tools/one_wmean.py 1e8 2 > gpurun_out/launches_wmean_r02.out 2>&1
timeout 900 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/launches_dtw_r02.csv \
    python tools/one_dtw.py 1e8 5 > gpurun_out/launches_dtw_r02.out 2>&1
ls -la gpurun_out
