mkdir -p gpurun_out
L=$PWD/build/libvp9b200
VP9B_LIB=${L}_aw.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_synthetic.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_aw.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_aw.log; tail -2 gpurun_out/pytest_aw.log
bash tools/ab.sh "cur|VP9B_LIB=${L}_cur.so|--steps 8" "aw|VP9B_LIB=${L}_aw.so|--steps 8" "cur2|VP9B_LIB=${L}_cur.so|--steps 8" "aw2|VP9B_LIB=${L}_aw.so|--steps 8" "aw-1440|VP9B_LIB=${L}_aw.so|--clip 1440p --streams 32" "aw-hbd|VP9B_LIB=${L}_aw.so|--clip hbd1080 --streams 32"
