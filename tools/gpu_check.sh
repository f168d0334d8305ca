#!/bin/bash
# One GPU-box call: parity tests, per-stage device times of the bench clips, a short bench line.
# Everything is wrapped in `timeout` so that a hung kernel cannot hold the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for clip in bench_1440p bench_4k; do
  timeout 300 python tools/replay.py $clip 1 --stages > gpurun_out/stages_$clip.txt 2>&1
  echo "stages $clip rc=$?"
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_quick.json
