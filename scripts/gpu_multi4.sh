N=${N:-4}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 12 --warmup 3 > gpurun_out/r2l_bench_${N}gpu.json 2> gpurun_out/r2l_bench_${N}gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2l_bench_${N}gpu.json; tail -5 gpurun_out/r2l_bench_${N}gpu.err
