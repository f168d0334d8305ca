mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_synthetic.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
bash tools/ab.sh "cpre||" "cpre-1440||--clip 1440p --streams 32"
BENCH="python bench.py --no-cpu --no-e2e --no-others --steps 1 --warmup 1 --passes 1 --batch 32"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^k_inter_pred_sb$ -s 40 -c 1 -f -o gpurun_out/r02b_full_k_inter_pred_sb $BENCH > gpurun_out/r02b_ncu_k_inter_pred_sb.log 2>&1
echo "mc rc=$?"
cp cuda-vp9_b200/libvp9b200.so gpurun_out/lib_for_prof_r02c.so
