mkdir -p gpurun_out/pv
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
bash tools/ab.sh "base|VP9B_LIB=build/libvp9b200_base.so|" "new||" "base2|VP9B_LIB=build/libvp9b200_base.so|" "new2||"
for rep in 1 2 3; do for n in v3 v5; do build/pv/$n/parse_bench tests/golden/clips/bench_4k.ivf 8 | awk -v n=$n '{print n, $4}'; done; done > gpurun_out/pv/ab2.txt; cat gpurun_out/pv/ab2.txt
