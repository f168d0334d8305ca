mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_synthetic.py tests/test_ref_binding.py -m gpu -x -q --timeout 120 > gpurun_out/pytest_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_last.log; tail -3 gpurun_out/pytest_last.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
