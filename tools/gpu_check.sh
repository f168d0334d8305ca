#!/bin/bash
# One GPU-box call: parity tests, per-stage device times of the bench clips, a short bench line,
# optional loop-filter phase stamps and one ncu capture.  Everything is wrapped in `timeout` so that
# a hung kernel cannot hold the box.   usage: gpu_check.sh [ncu-kernel-regex]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for clip in bench_1440p bench_4k; do
  timeout 300 python tools/replay.py $clip 1 --stages > gpurun_out/stages_$clip.txt 2>&1
  echo "stages $clip rc=$?"
done
if [ -f build/libvp9b200_timing.so ]; then
  VP9B_LIB=build/libvp9b200_timing.so timeout 120 python tools/lf_phases.py bench_1440p > gpurun_out/lf_phases.txt 2>&1
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_quick.json
if [ -n "$1" ]; then
  timeout 300 python tools/replay.py bench_1440p 1 > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 6 -c 2 -f -o gpurun_out/prof_$1 python tools/replay.py bench_1440p 1 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu rc=$?"
  cp cuda-vp9_b200/libvp9b200.so gpurun_out/lib_for_prof_$1.so   # the by-line tool needs the exact binary
fi
