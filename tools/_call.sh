mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-others > gpurun_out/r02_bench_4k_2gpu.json 2> gpurun_out/r02_bench_4k_2gpu.err
echo "rc=$?"; tail -c 400 gpurun_out/r02_bench_4k_2gpu.json
