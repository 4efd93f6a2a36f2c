N=${N:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2n_bench_${N}gpu.json 2> gpurun_out/r2n_bench_${N}gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2n_bench_${N}gpu.json; tail -4 gpurun_out/r2n_bench_${N}gpu.err
