# Sample-net diagnostics on the GPU box (through gpurun): parity of the tensor-core path, kernel times of the bench
# workload, and the clock64 / globaltimer timeline of CTA 0 (tile HR_TC_TRACE_ITER, default 1) printed by the library.
Q="python scripts/sweep.py --only technicolor_S32_K12 --rays 65536 --steps 10 --out gpurun_out/q.json"
timeout 600 python -m pytest tests/test_sample_net_tc_gpu.py -x -q 2>&1 | tail -3
echo "== kernel times"; $Q 2>&1 | tail -1
echo "== trace"; HR_TC_TRACE=1 $Q 2>&1 | grep "tc2-trace" | tail -17
