# ncu evidence for the bench command (run on the GPU box through gpurun; outputs under gpurun_out/).  TAG names the round.
TAG=${TAG:-r2a}
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"render_kernel|mlp_tc2_kernel" -s 4 -c 2 -f -o gpurun_out/${TAG}_prof $B > gpurun_out/${TAG}_prof.log 2>&1
ls -la gpurun_out/ | tail -5
