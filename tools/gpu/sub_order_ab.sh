# Morton sub-cell order inside a cell (DC_SUB_ORDER=1) against the caller's order (0): kNN + fixed-graph step times
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x 2>&1 | tail -3
for so in 0 1; do
  echo "== DC_SUB_ORDER=$so"
  DC_SUB_ORDER=$so python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-strong-anchor > gpurun_out/b_sub$so.json 2> gpurun_out/b_sub$so.err
  python - <<PY
import json
l = json.loads(open("gpurun_out/b_sub$so.json").read().strip().splitlines()[-1])
print({k: round(l[k], 3) for k in ("ms_per_step","search_ms","first_step_on_new_graph_ms","fixed_graph_step_ms","cold_search_ms")}, l["e2e"]["ms_per_step"])
print(l["entry_points_ms_per_step"])
PY
done
