#!/bin/bash
# 1 GPU, last check of the round: smoke + the in-situ engine test + plumbing with the final kernels.
O=gpurun_out/r2c21; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 400 python -m pytest tests/test_gpu_insitu.py tests/test_gpu_plumbing.py -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
