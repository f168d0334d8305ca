mkdir -p gpurun_out
L=$PWD/build/libvp9b200
VP9B_MC_PIPE=1 VP9B_LIB=${L}_pipe.so timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_synthetic.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pipe.log; tail -4 gpurun_out/pytest_pipe.log
bash tools/ab.sh "sb|VP9B_LIB=${L}_pipe.so|--steps 8" "pipe|VP9B_MC_PIPE=1 VP9B_LIB=${L}_pipe.so|--steps 8" "pipe-1440|VP9B_MC_PIPE=1 VP9B_LIB=${L}_pipe.so|--clip 1440p --streams 32" "pipe-hbd|VP9B_MC_PIPE=1 VP9B_LIB=${L}_pipe.so|--clip hbd1080 --streams 32"
