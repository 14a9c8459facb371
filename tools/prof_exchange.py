"""Where the slab plan + halo exchange spend their time (torchrun, one rank per GPU; developer tool).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/prof_exchange.py [n_scans_per_rank]
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, '.')


def main():
    rank, world, lr = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(lr)
    dev = torch.device('cuda', lr)
    os.environ.setdefault('NCCL_DEBUG', 'NONE')
    dist.init_process_group('nccl', device_id=dev)
    import depth_correction_b200 as dc
    from depth_correction_b200 import _lib as L
    from depth_correction_b200.preproc import _initial_map_points
    from bench import host_scans, make_poses, NN_R
    n_scans = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    ids = list(range(rank * n_scans, (rank + 1) * n_scans))
    pts, _ = host_scans(n_scans, 'os0-128', scene='corridor', scan_ids=ids)
    poses = torch.as_tensor(make_poses('corridor', n_scans * world), device=dev)
    clouds = []
    for p in pts:
        c = dc.DepthCloud.from_points(torch.as_tensor(p, device=dev))
        c.inc_angles = torch.rand((len(c), 1), device=dev)
        c.mask = torch.ones(len(c), dtype=torch.bool, device=dev)
        clouds.append(c)
    part = dc.SlabPartitioner()

    def once(profile):
        L.profile = {} if profile else None
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        wp = _initial_map_points(dc.global_cloud(clouds=clouds, poses=poses[ids]))
        torch.cuda.synchronize(); t1 = time.perf_counter()
        axis, bounds = part.plan(wp)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        loc = part.exchange(clouds, ids, wp, axis, bounds, halo=NN_R)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        prof = L.collect_profile() if profile else {}
        L.profile = None
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, prof, loc

    for _ in range(3):
        once(False)
    a, b, c, prof, loc = once(True)
    if rank == 0:
        print('world points %.2f ms | plan %.2f ms | exchange %.2f ms | local points %d' % (a, b, c, len(loc)))
        for k, v in sorted(prof.items()):
            print('   %-28s %.3f ms' % (k, v['ms_total']))
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as pr:
            once(False)
        if rank == 0:
            print(pr.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
    except Exception as e:      # noqa: BLE001
        if rank == 0:
            print('torch profiler unavailable:', e)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
