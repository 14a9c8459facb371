#!/bin/bash
# cell-size target of the kNN search (points per occupied cell = DC_KNN_OCC * k) with the recorded kernel, bench map
cd "$GRAFT_REPO_ROOT"
for occ in 0.15 0.2 0.3 0.45 0.6 0.9; do
  echo "DC_KNN_OCC=$occ"
  DC_KNN_OCC=$occ timeout 300 python tools/check_knn_recorded.py 64 2>&1 | grep "64 scans\|street"
done
