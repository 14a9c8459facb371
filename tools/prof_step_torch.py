"""torch.profiler view of one search + fused step of the bench workload (developer tool): every kernel of the timed region
incl. torch's own (memsets, copies, elementwise glue) and the host time per step.

    python tools/prof_step_torch.py [n_scans]
"""
import sys
import time

import torch

sys.path.insert(0, '.')
import depth_correction_b200 as dc                      # noqa: E402
from depth_correction_b200 import _lib as L             # noqa: E402
from bench import Job                                   # noqa: E402
from torch.profiler import profile, ProfilerActivity    # noqa: E402

n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device('cuda:0')
job = Job(dc, dev, 1, 0, 'corridor', range(n_scans), n_scans)
for _ in range(3):
    job.step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    job.step()
host = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 5 * 1e3
print('host time per step %.2f ms, wall %.2f ms' % (host, wall))
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
    for _ in range(2):
        job.step()
    torch.cuda.synchronize()
print(pr.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=70))
