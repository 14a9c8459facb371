"""Worker of tests/test_gpu_multi.py (launched with torchrun, one rank per GPU, NCCL): the sharded fused
step (spatial slabs + halo + one all-reduce) must reproduce the single-GPU fused step on the same map."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local_rank = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    import depth_correction_b200 as dc
    from depth_correction_b200.synthetic import make_sequence

    S = 4 * world
    scans_np, _, poses_np = make_sequence('corridor', n_scans=S, pattern='os0-32', seed=3, grid_res=0.15, step=1.5)
    results = {}
    for k, r in ((0, 0.4), (12, 0.4)):
        cfg = dc.Config(nn_k=k, nn_r=r, pose_correction=dc.PoseCorrection.pose)
        clouds = [dc.local_feature_cloud(dc.DepthCloud.from_points(torch.as_tensor(s['points'], device=dev)), cfg)
                  for s in scans_np]
        clouds = [dc.DepthCloud(vps=c.vps, dirs=c.dirs, depth=c.depth, inc_angles=c.inc_angles, mask=c.mask) for c in clouds]
        poses = torch.as_tensor(poses_np, device=dev)
        d0 = torch.as_tensor(np.random.default_rng(0).normal(0, 0.005, (S, 6)), device=dev)

        def run(local_clouds, local, inlier_ratio=1.0):
            model = dc.ScaledPolynomial(w=[0.003, -0.002], exponent=[2, 4], device=dev)
            deltas = d0.clone().requires_grad_(True)
            sel = None if local is None else local.scan_ids
            ns = dc.establish_neighborhoods(clouds=local_clouds, poses=poses if sel is None else poses[sel], cfg=cfg)
            pc = torch.stack(dc.create_corrected_poses(poses, deltas, cfg))
            feats = dc.compute_neighborhood_features(
                cloud=dc.global_cloud(clouds=local_clouds, model=model, poses=pc if sel is None else pc[sel]),
                neighborhoods=ns, cfg=cfg)
            if local is None:
                loss, _ = dc.trace_loss(feats, sqrt=True, inlier_ratio=inlier_ratio)
                loss.backward()
            elif inlier_ratio < 1.0:
                # global inlier threshold over the owned points of all ranks (distributed quantile)
                sc = dc.sharded_inlier_sum_count(feats, local.owned, inlier_ratio=inlier_ratio, loss='trace_loss', sqrt=True)
                loss = dc.reduce_step(sc, [model.w, deltas])
            else:
                sc = dc.fused_sum_count(feats, mask=local.owned, loss='trace_loss', sqrt=True)
                loss = dc.reduce_step(sc, [model.w, deltas])
            return loss.detach(), model.w.grad.clone(), deltas.grad.clone()

        ref = run(clouds, None)                                   # every rank: the whole map on its own GPU
        mine = list(range(rank, S, world))                        # interleaved ingestion: forces a real exchange
        wp = [clouds[i].transform(poses[i]).to_points() for i in mine]
        part = dc.SlabPartitioner()
        axis, bounds = part.plan(wp)
        local = part.exchange([clouds[i] for i in mine], mine, wp, axis, bounds, halo=r)
        got = run(local.clouds, local)
        for a, b, name in zip(got, ref, ('loss', 'w_grad', 'pose_grad')):
            err = (a - b).abs().max().item() / b.abs().max().item()
            results['%s k=%d' % (name, k)] = err
            assert err < 1e-9, (name, k, err)
        ref_in = run(clouds, None, inlier_ratio=0.7)
        got_in = run(local.clouds, local, inlier_ratio=0.7)
        for a, b, name in zip(got_in, ref_in, ('loss', 'w_grad', 'pose_grad')):
            err = (a - b).abs().max().item() / b.abs().max().item()
            results['inlier %s k=%d' % (name, k)] = err
            assert err < 1e-9, ('inlier', name, k, err)
    if rank == 0:
        print('MULTIGPU_OK', results)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
