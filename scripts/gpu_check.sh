#!/bin/bash
# One GPU-box visit: the -m gpu suites in two batches (a hang in the tensor-core suites must not take the rest with it),
# then smoke().  Logs land in gpurun_out/.
mkdir -p gpurun_out
T="--timeout 300 --timeout-method thread"
timeout 1200 python -m pytest tests -m gpu -q $T --deselect tests/test_sample_net_tc_gpu.py --deselect tests/test_shipped_yaml_gpu.py --deselect tests/test_widened_gpu.py -k "not auto and not tensor_core" > gpurun_out/pytest_a.log 2>&1; echo "batch A rc=$?" 
tail -15 gpurun_out/pytest_a.log
timeout 1200 python -m pytest tests/test_sample_net_tc_gpu.py tests/test_shipped_yaml_gpu.py tests/test_parity_bites_gpu.py tests/test_widened_gpu.py -m gpu -q $T > gpurun_out/pytest_b.log 2>&1; echo "batch B rc=$?"
tail -15 gpurun_out/pytest_b.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
