# multi-GPU checks on an N-GPU box: NCCL parity test + bench line (N = $1)
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/t_multi_n$N.log 2>&1; tail -5 gpurun_out/t_multi_n$N.log
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/b_r2_n$N.json 2> gpurun_out/b_r2_n$N.err ) 2> gpurun_out/b_r2_n$N.time
tail -5 gpurun_out/b_r2_n$N.err; cat gpurun_out/b_r2_n$N.time
python - <<PY
import json
l = json.loads(open("gpurun_out/b_r2_n$N.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","scaling","search_ms","first_step_on_new_graph_ms","cold_search_ms","fixed_graph_step_ms","setup_ms","e2e","weak","strong_scaling","multi_gpu_parity","grad_check","gpu_launches"):
    print(k, l.get(k))
print(l["config"])
PY
