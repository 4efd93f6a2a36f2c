Q="python scripts/sweep.py --only technicolor_S32_K12 --rays 65536 --steps 10 --out gpurun_out/q.json"
echo "== trace tile 0"; HR_TC_TRACE_ITER=0 HR_TC_TRACE=1 $Q 2>&1 | grep "tc2-trace" | tail -17
