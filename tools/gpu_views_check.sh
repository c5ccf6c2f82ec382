#!/bin/bash
# Lean GPU-box visit for the view pipeline only: parity tests of every W-axis kernel + the kernel / tile-height sweep.
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_views.py -q > gpurun_out/views_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/views_pytest.log
timeout 40 python tools/views_timing.py 4 > gpurun_out/views_timing2.jsonl 2>&1; echo "timing rc=$?"; cut -c1-150 gpurun_out/views_timing2.jsonl
