mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
bash tools/ab.sh "ring||" "ring-1440||--clip 1440p --streams 32" "ring-hbd||--clip hbd1080 --streams 32"
