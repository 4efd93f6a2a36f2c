mkdir -p gpurun_out
# ncu reports are ~40 MB each and gpurun_out is capped at 64 MiB: summarise on the box, keep only the summaries
for w in donerf_sphere_s16 neural3d_s64; do
  timeout 600 ncu --set full --clock-control none -k regex:"render_kernel" -s 2 -c 1 -f -o /tmp/r2g_$w python scripts/run_workload.py $w 4 > gpurun_out/r2g_$w.log 2>&1; echo "ncu $w rc=$?"
  python scripts/ncu_summary.py /tmp/r2g_$w.ncu-rep r2g_$w gpurun_out
done
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches.csv $B > gpurun_out/r2g_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"render_kernel|mlp_tc2_kernel" -s 4 -c 2 -f -o /tmp/r2g_prof $B > gpurun_out/r2g_prof.log 2>&1; echo "prof rc=$?"
python scripts/ncu_summary.py /tmp/r2g_prof.ncu-rep r2g gpurun_out
ls -la gpurun_out | tail -12
