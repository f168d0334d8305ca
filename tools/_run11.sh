python tools/replay.py bench_1440p 1 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_inter_pred -s 6 -c 2 -o gpurun_out/prof_mc2 python tools/replay.py bench_1440p 1 > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
ncu --set full --clock-control none --import-source on -k regex:k_residual -s 6 -c 2 -o gpurun_out/prof_res python tools/replay.py bench_1440p 1 > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
ncu --set full --clock-control none --import-source on -k regex:k_intra -s 6 -c 2 -o gpurun_out/prof_intra python tools/replay.py bench_1440p 1 > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log
