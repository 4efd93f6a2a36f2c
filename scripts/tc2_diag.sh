Q="python scripts/sweep.py --only technicolor_S32_K12 --rays 65536 --steps 10 --out gpurun_out/q.json"
timeout 600 python -m pytest tests/test_sample_net_tc_gpu.py tests/test_parity_gpu.py -x -q 2>&1 | tail -4
echo "== normal"; $Q 2>&1 | tail -1
echo "== trace"; HR_TC_TRACE=1 $Q 2>&1 | grep "tc2-trace" | tail -17
