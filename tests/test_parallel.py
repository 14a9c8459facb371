"""Multi-process (gloo, world_size 2, CPU) test of the spatial-slab partition, halo exchange and the
one-buffer all-reduce of depth_correction_b200.parallel.  The per-rank arithmetic is done by the CPU oracle
(the kernels need a GPU); what is under test is the host-side logic: every loss term is owned exactly once,
owned points see all their neighbours through the halo, and loss / gradients of the sharded run equal the
single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

R = 0.4


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dataset():
    from depth_correction_b200.synthetic import make_sequence
    from oracle import oracle
    scans_np, _, poses = make_sequence('corridor', n_scans=6, pattern='os0-32', seed=9, grid_res=0.25, step=2.0)
    rng = np.random.default_rng(4)
    scans = []
    for sc in scans_np:
        pts = torch.as_tensor(sc['points'].astype(np.float64))
        vps, dirs, depth = oracle.from_points(pts)
        n = len(pts)
        scans.append({'vps': vps, 'dirs': dirs, 'depth': depth,
                      'inc_angles': torch.as_tensor(rng.uniform(0.0, 1.4, (n, 1))),
                      'mask': torch.as_tensor(rng.uniform(size=n) < 0.8)})
    return scans, torch.as_tensor(poses)


def _local_sum_count(oracle, scans, poses, w, e, owned):
    pts, _ = oracle.global_points(scans, poses, w, e, True)
    _, nb = oracle.nearest_neighbors(pts.detach(), r=R)
    f = oracle.neighborhood_features(pts, nb, eigvecs=False)
    loss, _ = oracle.min_eigval_loss(f['eigvals'], owned, normalization=True, reduction='sum')
    return torch.stack([loss, owned.sum().to(loss.dtype)]), nb


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import depth_correction_b200 as dc
        from oracle import oracle
        scans, poses = _dataset()
        S = len(scans)
        w0 = torch.tensor([[0.004, -0.002]], dtype=torch.float64)
        e = torch.tensor([[2.0, 4.0]], dtype=torch.float64)
        d0 = torch.as_tensor(np.random.default_rng(1).normal(0, 0.01, (S, 6)))

        # ---- single-process reference on the whole map
        w = w0.clone().requires_grad_(True)
        d = d0.clone().requires_grad_(True)
        full, nb_full = _local_sum_count(oracle, scans, oracle.create_corrected_poses(poses, d), w, e,
                                         torch.ones(sum(len(s['depth']) for s in scans), dtype=torch.bool))
        (full[0] / full[1]).backward()
        ref_loss, ref_gw, ref_gd = (full[0] / full[1]).item(), w.grad.clone(), d.grad.clone()

        # ---- sharded run: rank ingests scans rank::world
        mine = list(range(rank, S, world))
        clouds = [dc.DepthCloud(vps=scans[i]['vps'], dirs=scans[i]['dirs'], depth=scans[i]['depth'],
                                inc_angles=scans[i]['inc_angles'], mask=scans[i]['mask']) for i in mine]
        world_pts = [c.transform(poses[i]).to_points() for c, i in zip(clouds, mine)]
        part = dc.SlabPartitioner()
        axis, bounds = part.plan(world_pts)
        local = part.exchange(clouds, mine, world_pts, axis, bounds, halo=R)
        assert axis == 0 and len(bounds) == world + 1
        # every point is owned by exactly one rank
        gathered = [None] * world
        dist.all_gather_object(gathered, local.global_ids[local.owned].tolist())
        owned_all = sorted(tuple(x) for part_ in gathered for x in part_)
        expect = sorted((s, i) for s in range(S) for i in range(len(scans[s]['depth'])))
        assert owned_all == expect
        # local set == all points within `halo` of the slab, records intact
        all_pts, _ = oracle.global_points(scans, poses)
        offs = np.cumsum([0] + [len(s['depth']) for s in scans])
        x = all_pts[:, axis]
        member = (x >= local.bounds[0] - R) & (x < local.bounds[1] + R)
        gid = (offs[local.global_ids[:, 0].numpy()] + local.global_ids[:, 1].numpy())
        assert sorted(gid.tolist()) == member.nonzero().squeeze(1).tolist()
        assert torch.equal(local.owned, ((x >= local.bounds[0]) & (x < local.bounds[1]))[gid])
        lscans = [{'vps': c.vps, 'dirs': c.dirs, 'depth': c.depth, 'inc_angles': c.inc_angles, 'mask': c.mask}
                  for c in local.clouds]
        cat = lambda k: torch.cat([s[k] for s in scans])
        assert torch.equal(torch.cat([c.depth for c in local.clouds]), cat('depth')[gid])
        assert torch.equal(torch.cat([c.mask for c in local.clouds]), cat('mask')[gid])
        # neighbourhoods of owned points are complete: same global neighbour sets as the single-process graph
        w = w0.clone().requires_grad_(True)
        d = d0.clone().requires_grad_(True)
        poses_c = oracle.create_corrected_poses(poses, d)[local.scan_ids]
        sc, nb_loc = _local_sum_count(oracle, lscans, poses_c, w, e, local.owned)
        g_of_l = torch.as_tensor(gid)
        nb_glob = torch.where(nb_loc >= 0, g_of_l[nb_loc.clamp(min=0)], nb_loc)
        for row in local.owned.nonzero().squeeze(1)[::37].tolist():
            assert sorted(v for v in nb_glob[row].tolist() if v >= 0) == sorted(v for v in nb_full[gid[row]].tolist() if v >= 0)
        # one all-reduce: loss and gradients equal the single-process run
        loss = dc.reduce_step(sc, [w, d])
        assert abs(loss.item() - ref_loss) < 1e-12 * abs(ref_loss)
        assert (w.grad - ref_gw).abs().max() < 1e-10 * ref_gw.abs().max()
        assert (d.grad - ref_gd).abs().max() < 1e-10 * ref_gd.abs().max()
        out.put((rank, 'ok', len(local), int(local.owned.sum())))
    except Exception as ex:   # surface the failure in the parent
        import traceback
        out.put((rank, 'fail', traceback.format_exc(), repr(ex)))
    finally:
        dist.destroy_process_group()


def test_slab_partition_halo_and_allreduce_world2():
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert r[1] == 'ok', r[2]
    n_local = sum(r[2] for r in results)
    n_owned = sum(r[3] for r in results)
    assert n_local > n_owned > 0          # halo copies exist


def test_single_process_partition_is_identity():
    sys.path.insert(0, ROOT)
    import depth_correction_b200 as dc
    scans, poses = _dataset()
    clouds = [dc.DepthCloud(vps=s['vps'], dirs=s['dirs'], depth=s['depth'], inc_angles=s['inc_angles'], mask=s['mask'])
              for s in scans]
    world_pts = [c.transform(poses[i]).to_points() for i, c in enumerate(clouds)]
    part = dc.SlabPartitioner()
    axis, bounds = part.plan(world_pts)
    local = part.exchange(clouds, list(range(len(clouds))), world_pts, axis, bounds, halo=R)
    assert local.owned.all() and len(local) == sum(len(c) for c in clouds)
    assert all(torch.equal(a.depth, b.depth) for a, b in zip(local.clouds, clouds))


def _quantile_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import depth_correction_b200 as dc
        rng = np.random.default_rng(5)
        cases = {
            'lognormal': rng.lognormal(-9.0, 2.0, 300001),                 # loss-like: spans many decades
            'ties': np.round(rng.random(50000), 2),                        # 101 distinct values
            'tiny': rng.random(7),
            'constant': np.full(1000, 0.25),
            'outlier': np.concatenate([rng.random(20000) * 1e-6, [1e9]]),  # one value stretches the range
        }
        for name, full in cases.items():
            shard = torch.as_tensor(full[rank::world] if name != 'tiny' else (full if rank == 0 else full[:0]))
            for q in (0.0, 0.3, 0.5, 0.9, 0.999, 1.0):
                ours = dc.distributed_quantile(shard, q, max_gather=64).item()
                ref = torch.quantile(torch.as_tensor(full), q).item()
                assert ours == ref, (name, q, ours, ref)
        out.put((rank, 'ok'))
    except Exception as ex:
        import traceback
        out.put((rank, 'fail', traceback.format_exc(), repr(ex)))
    finally:
        dist.destroy_process_group()


def test_distributed_quantile_equals_torch_quantile_world2():
    """The global inlier threshold of a sharded map (inlier_ratio < 1, loss.py:256-267): bit-equal to
    torch.quantile of the concatenated shards, without gathering them."""
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_quantile_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in results:
        assert r[1] == 'ok', r[2]
