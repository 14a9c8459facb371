#!/bin/bash
# captured-iteration test, the small-map host profile, and the small-config bench lines
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q -k "captured" 2>&1 | tail -15
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-strong-anchor 2>gpurun_out/cap_bench.err | python -c "
import json,sys
l=json.loads(sys.stdin.readlines()[-1])
print(json.dumps(l['other_configs'], indent=1))
print(l['value'], l['ms_per_step'], l['e2e'])
"
tail -5 gpurun_out/cap_bench.err
