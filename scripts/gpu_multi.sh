# multi-GPU visit (gpurun --gpus N): bit-identity of ray-sharded rendering over NCCL / peer memory, then the bench at N ranks
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_parity_bites_gpu.py -m gpu -q -k sharded --timeout 600 --timeout-method thread > gpurun_out/pytest_multi.log 2>&1; echo "sharded test rc=$?"; tail -20 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2k_bench_${N}gpu.json 2> gpurun_out/r2k_bench_${N}gpu.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2k_bench_${N}gpu.json; tail -15 gpurun_out/r2k_bench_${N}gpu.err
