mkdir -p gpurun_out
L=$PWD/build/libvp9b200
VP9B_LIB=${L}_v5.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_synthetic.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_v5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v5.log; tail -2 gpurun_out/pytest_v5.log
bash tools/ab.sh "v2b|VP9B_LIB=${L}_v2b.so|--steps 6" "v5|VP9B_LIB=${L}_v5.so|--steps 6" "v2b-2|VP9B_LIB=${L}_v2b.so|--steps 6" "v5-2|VP9B_LIB=${L}_v5.so|--steps 6" "v2b-1440|VP9B_LIB=${L}_v2b.so|--clip 1440p --streams 32" "v5-1440|VP9B_LIB=${L}_v5.so|--clip 1440p --streams 32" "v2b-hbd|VP9B_LIB=${L}_v2b.so|--clip hbd1080 --streams 32"  "v5-hbd|VP9B_LIB=${L}_v5.so|--clip hbd1080 --streams 32"
