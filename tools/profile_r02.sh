#!/bin/bash
# ncu evidence of round 2 (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are made here by
# tools/summarize_profiles.py).  Numbers printed by programs under ncu are never bench values.
set -u
mkdir -p gpurun_out /tmp/prof
R=/tmp/prof   # .ncu-rep files stay on the box (gpurun_out is capped at 64 MiB); their summaries and two reports come back
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
# 1. launch list of the bench command itself (headline step, n = 1e9)
timeout 900 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 10 --warmup 3 --no-side > gpurun_out/launches_r02.out 2>&1
# 2. the same step at BASELINE configs[1] (n = 1e8)
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_r02.csv \
    python bench.py --n 1e8 --steps 10 --warmup 3 --no-side > gpurun_out/launches_cfg2_r02.out 2>&1
# 3. --set full of the dominant kernel at both sizes (a selective query), and at n = 1e8 on a query with 10 M gate passes
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cnsm_stream -s 3 -c 1 -f -o $R/stream_1e8_r02 \
    python tools/one_query.py 1e8 1024 2048 5 5 > gpurun_out/stream_1e8_r02.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cnsm_stream -s 3 -c 1 -f -o $R/stream_dense_1e8_r02 \
    python tools/one_query.py 1e8 1024 2048 5 5 2736494 > gpurun_out/stream_dense_1e8_r02.out 2>&1
timeout 900 ncu --set full --clock-control none -k regex:cnsm_stream -s 2 -c 1 -f -o $R/stream_1e9_r02 \
    python tools/one_query.py 1e9 1024 2048 5 4 > gpurun_out/stream_1e9_r02.out 2>&1
# 4. the tail kernels of a cNSM-ED query, the window-mean pass, the cNSM-DTW cascade on a heavy query, RSM-ED
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chain_rewalk_kernel|cnsm_ed_exact|cnsm_ed_screen" -s 6 -c 3 -f -o $R/tail_1e8_r02 \
    python tools/one_query.py 1e8 1024 2048 5 5 > gpurun_out/tail_1e8_r02.out 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'wmean|rewalk_batch' -s 2 -c 2 -f -o $R/wmean_1e8_r02 \
    python tools/one_wmean.py 1e8 2 > gpurun_out/wmean_1e8_r02.out 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/launches_wmean_r02.csv \
    python tools/one_wmean.py 1e8 2 > gpurun_out/launches_wmean_r02.out 2>&1
timeout 900 ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/launches_dtw_r02.csv \
    python tools/one_dtw.py 1e8 5 6154940 2 > gpurun_out/launches_dtw_r02.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dtw_probe|dtw_lb_fused|dtw_band_coop" -s 3 -c 3 -f -o $R/dtw_heavy_r02 \
    python tools/one_dtw.py 1e8 5 6154940 2 > gpurun_out/dtw_heavy_r02.out 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ed_verify -s 2 -c 1 -f -o $R/rsm_ed_1e8_r02 \
    python tools/rsm_bench.py 1e8 > gpurun_out/rsm_ed_1e8_r02.out 2>&1
S="python tools/summarize_profiles.py"
$S launches gpurun_out/launches_r02.csv gpurun_out/launches_r02.md "round 2, \`python bench.py --steps 10 --warmup 3 --no-side\` (n = 1e9, m = 1024, chains of 2048)" > /dev/null
$S launches gpurun_out/launches_cfg2_r02.csv gpurun_out/launches_cfg2_r02.md "round 2, BASELINE configs[1]: \`python bench.py --n 1e8 --steps 10 --warmup 3 --no-side\`" > /dev/null
$S launches gpurun_out/launches_wmean_r02.csv gpurun_out/launches_wmean_r02.md "round 2, fused window-mean pass: \`python tools/one_wmean.py 1e8 2\`" > /dev/null
$S launches gpurun_out/launches_dtw_r02.csv gpurun_out/launches_dtw_r02.md "round 2, cNSM-DTW heavy query: \`python tools/one_dtw.py 1e8 5 6154940 2\` (m = 2048, rho = 102, eps = 5; 2.0 M flagged windows)" > /dev/null
$S full $R/stream_1e8_r02.ncu-rep gpurun_out/stream_1e8_r02.md "cnsm_stream_kernel, n = 1e8, m = 1024, selective query (tools/one_query.py 1e8 1024 2048 5 5)" > /dev/null
$S full $R/stream_dense_1e8_r02.ncu-rep gpurun_out/stream_dense_1e8_r02.md "cnsm_stream_kernel, n = 1e8, m = 1024, query with 10.5 M gate passes (tools/one_query.py 1e8 1024 2048 5 5 2736494)" > /dev/null
$S full $R/stream_1e9_r02.ncu-rep gpurun_out/stream_1e9_r02.md "cnsm_stream_kernel, n = 1e9, m = 1024, selective query (tools/one_query.py 1e9 1024 2048 5 4)" > /dev/null
$S full $R/tail_1e8_r02.ncu-rep gpurun_out/tail_1e8_r02.md "chain_rewalk_kernel, cnsm_ed_screen_kernel, cnsm_ed_exact_kernel, n = 1e8 (tools/one_query.py 1e8 1024 2048 5 5)" > /dev/null
$S full $R/wmean_1e8_r02.ncu-rep gpurun_out/wmean_1e8_r02.md "fused window-mean pass, n = 1e8, five widths (tools/one_wmean.py 1e8 2)" > /dev/null
$S full $R/dtw_heavy_r02.ncu-rep gpurun_out/dtw_heavy_r02.md "cNSM-DTW heavy query (n = 1e8, m = 2048, rho = 102, eps = 5, query 6154940: 2.0 M flagged windows, 0.8 M LB survivors)" > /dev/null
$S full $R/rsm_ed_1e8_r02.ncu-rep gpurun_out/rsm_ed_1e8_r02.md "ed_verify_kernel, n = 1e8, m = 1024, eps = 10 (tools/rsm_bench.py 1e8)" > /dev/null
cp $R/stream_dense_1e8_r02.ncu-rep $R/dtw_heavy_r02.ncu-rep gpurun_out/ 2>/dev/null
ls -la gpurun_out | head -60; du -sh gpurun_out
