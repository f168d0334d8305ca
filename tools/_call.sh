mkdir -p gpurun_out
L=$PWD/build/libvp9b200
VP9B_LIB=${L}_iq.so timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_synthetic.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_iq.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iq.log; tail -4 gpurun_out/pytest_iq.log
bash tools/ab.sh "iq|VP9B_LIB=${L}_iq.so|--steps 8" "head||--steps 8" "iq-hbd|VP9B_LIB=${L}_iq.so|--clip hbd1080 --streams 32"
