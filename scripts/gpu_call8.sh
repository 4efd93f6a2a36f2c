mkdir -p gpurun_out
# final records of the round with the final binary: sanitizer (memcheck + racecheck) over every kernel family, ncu launch list
# and full captures summarised on the box (reports are too large to travel), the bench line
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize.py > gpurun_out/r2j_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2j_sanitizer_memcheck.log
HR_SANITIZE_RAYS=200 timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize.py > gpurun_out/r2j_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2j_sanitizer_racecheck.log
for w in donerf_sphere_s16 neural3d_s64; do
  timeout 600 ncu --set full --clock-control none -k regex:"render_kernel" -s 2 -c 1 -f -o /tmp/r2j_$w python scripts/run_workload.py $w 4 > gpurun_out/r2j_$w.log 2>&1; echo "ncu $w rc=$?"
  python scripts/ncu_summary.py /tmp/r2j_$w.ncu-rep r2j_$w gpurun_out
done
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv $B > gpurun_out/r2j_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"render_kernel|mlp_tc2_kernel" -s 4 -c 2 -f -o /tmp/r2j_prof $B > gpurun_out/r2j_prof.log 2>&1; echo "prof rc=$?"
python scripts/ncu_summary.py /tmp/r2j_prof.ncu-rep r2j gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench_1gpu.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2j_bench.err
ls -la gpurun_out | tail -14
