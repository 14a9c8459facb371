#!/bin/bash
# the bench map against the CPU oracle at full size + the whole GPU test suite
cd "$GRAFT_REPO_ROOT"
free -g | head -2; nproc
timeout 300 python tools/full_map_parity.py --scans 4 --chunk 100000 2>&1 | tail -30
timeout 1500 python tools/full_map_parity.py > gpurun_out/full_map_parity.json 2> gpurun_out/full_map_parity.err; echo "rc=$?"
cat gpurun_out/full_map_parity.json; tail -3 gpurun_out/full_map_parity.err
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
