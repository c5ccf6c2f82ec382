#!/bin/bash
# One GPU-box visit: parity suite, view-pipeline bench + ncu capture, smoke, default bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 90 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_pytest.log
timeout 40 python tools/views_timing.py 4 > gpurun_out/views_timing.jsonl 2>&1; echo "timing rc=$?"; cat gpurun_out/views_timing.jsonl | cut -c1-160
timeout 40 python bench.py --workload views --steps 10 > gpurun_out/b_views.json 2> gpurun_out/b_views.err; echo "views rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/b_views.json").read().strip().splitlines()[-1])
    print("views", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["achieved"], d["roofline"]["frac"], "delta", d["score_delta_vs_oracle"], d["breakdown"])
except Exception as e:
    print("views parse failed", e)
PY
timeout 45 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/views_full python tools/views_profile.py 1 > gpurun_out/views_ncu.log 2>&1; echo "ncu rc=$?"
if [ "$1" = "full" ]; then
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 45 python bench.py --steps 10 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/final_bench.json
fi
