mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
( time timeout 900 python bench.py > gpurun_out/r02_bench_4k.json 2> gpurun_out/r02_bench_4k.err ) 2> gpurun_out/r02_bench_time.txt
echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench_4k.json
( time timeout 600 python bench.py --impl reference > gpurun_out/r02_ref_4k.json 2> gpurun_out/r02_ref_4k.err ) 2>> gpurun_out/r02_bench_time.txt
echo "ref rc=$?"; tail -c 300 gpurun_out/r02_ref_4k.json
for clip in bench_1440p bench_4k; do
  timeout 300 python tools/replay.py $clip 1 --stages > gpurun_out/r02_stages_${clip#bench_}.txt 2>&1
done
bash tools/profile_r02.sh
cp cuda-vp9_b200/libvp9b200.so gpurun_out/lib_for_prof_r02.so
