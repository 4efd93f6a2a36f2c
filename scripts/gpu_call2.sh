mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_1gpu.json 2> gpurun_out/r2a_bench_1gpu.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2a_bench_1gpu.json; tail -5 gpurun_out/r2a_bench_1gpu.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>&1; echo "ref rc=$?"; tail -c 800 gpurun_out/r2a_bench_ref.json
TAG=r2a bash scripts/profile.sh > gpurun_out/r2a_profile.log 2>&1; echo "profile rc=$?"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a_memcheck.log python scripts/sanitize.py > gpurun_out/r2a_memcheck.out 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2a_memcheck.out; tail -3 gpurun_out/r2a_memcheck.log
HR_SANITIZE_RAYS=300 timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r2a_racecheck.log python scripts/sanitize.py > gpurun_out/r2a_racecheck.out 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2a_racecheck.out; tail -3 gpurun_out/r2a_racecheck.log
timeout 300 python -m pytest tests/test_parity_bites_gpu.py -m gpu -q -k embed 2>&1 | tail -3
