#!/bin/bash
# Lean GPU-box visit for the view pipeline only: parity tests of every W-axis kernel, the bench line, and one ncu --set full
# capture (durations included) of the default path.
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_views.py -q > gpurun_out/views_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/views_pytest.log
timeout 40 python bench.py --workload views --steps 10 > gpurun_out/b_views.json 2> gpurun_out/b_views.err; echo "views rc=$?"; cut -c1-100 gpurun_out/b_views.json; grep -o '"roofline.*"cpu_baseline' gpurun_out/b_views.json | cut -c1-400
timeout 40 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:resize_ -c 3 -f -o gpurun_out/views_full2 python tools/views_profile.py 1 > gpurun_out/views_ncu2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/views_ncu2.log
