"""CPU tests of the kernels' per-point fp64 math (dc_math.cuh compiled for the host) and of the C ABI."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def hm():
    out = os.path.join(ROOT, 'tests', '_emu')
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, 'libhostmath.so')
    src = os.path.join(ROOT, 'tests', 'host_math', 'host_math.cpp')
    subprocess.check_call(['g++', '-O2', '-shared', '-fPIC', '-x', 'c++', src, '-o', so])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_eig3_against_lapack(hm):
    rng = np.random.default_rng(0)
    n = 20000
    # planar (lambda0 << lambda2), linear, isotropic and generic neighbourhood covariances
    scales = np.concatenate([np.tile([5e-3, 0.2, 0.25], (n // 4, 1)), np.tile([5e-3, 6e-3, 0.3], (n // 4, 1)),
                             np.tile([0.1, 0.1, 0.1], (n // 4, 1)), rng.uniform(1e-3, 1, (n // 4, 3))])
    X = rng.normal(0, 1, (n, 24, 3)) * scales[:, None, :]
    Q = np.linalg.qr(rng.normal(0, 1, (n, 3, 3)))[0]
    X = X @ Q + rng.uniform(-50, 50, (n, 1, 3))
    Xc = X - X.mean(1, keepdims=True)
    C = np.ascontiguousarray(np.einsum('nki,nkj->nij', Xc, Xc) / 23)
    lam = np.empty((n, 3))
    V = np.empty((n, 3, 3))
    hm.hm_eig(_p(C), ctypes.c_long(n), _p(lam), _p(V))
    ref = np.linalg.eigvalsh(C)
    assert np.max(np.abs(lam - ref) / np.abs(ref)) < 1e-9         # per-eigenvalue relative error
    assert np.max(np.abs(lam - ref) / ref[:, 2:3]) < 1e-14        # relative to the matrix norm
    assert np.max(np.abs(np.einsum('nij,nj,nkj->nik', V, lam, V) - C)) < 1e-11
    assert np.max(np.abs(np.einsum('nij,nik->njk', V, V) - np.eye(3))) < 1e-10
    assert np.all(np.diff(lam, axis=1) >= 0)


def test_eig3_special_cases(hm):
    mats = np.zeros((5, 3, 3))
    mats[1] = np.eye(3) * 2.5
    mats[2] = np.diag([1.0, 1.0, 3.0])
    mats[3] = np.diag([3.0, 0.0, 3.0])
    mats[4] = np.outer([1.0, 2.0, -1.0], [1.0, 2.0, -1.0])
    lam = np.empty((5, 3))
    V = np.empty((5, 3, 3))
    hm.hm_eig(_p(mats), ctypes.c_long(5), _p(lam), _p(V))
    assert np.all(np.isfinite(lam)) and np.all(np.isfinite(V))
    assert np.max(np.abs(lam - np.linalg.eigvalsh(mats))) < 1e-14
    assert np.max(np.abs(np.einsum('nij,nj,nkj->nik', V, lam, V) - mats)) < 1e-14


def test_pose_compose_and_reverse_mode(hm, golden):
    g = golden('misc')
    d = np.ascontiguousarray(g['xyz_axis_angle'])
    n = len(d)
    rng = np.random.default_rng(1)
    P = oracle.xyz_axis_angle_to_matrix(torch.as_tensor(rng.normal(0, 1, (n, 6)))).numpy().copy()
    T = np.empty((n, 12))
    hm.hm_pose(_p(P), _p(d), ctypes.c_long(n), _p(T))
    ref = oracle.create_corrected_poses(torch.as_tensor(P), torch.as_tensor(d)).numpy()
    assert np.max(np.abs(T.reshape(n, 3, 4) - ref[:, :3, :])) < 1e-14
    G = rng.normal(0, 1, (n, 12))
    for deltas in (d, np.zeros_like(d)):
        dt = torch.as_tensor(deltas).clone().requires_grad_(True)
        (oracle.create_corrected_poses(torch.as_tensor(P), dt)[:, :3, :] * torch.as_tensor(G.reshape(n, 3, 4))).sum().backward()
        gd = np.empty((n, 6))
        hm.hm_pose_bwd(_p(P), _p(np.ascontiguousarray(deltas)), _p(G), ctypes.c_long(n), _p(gd))
        assert np.max(np.abs(gd - dt.grad.numpy())) < 1e-12


def test_pow_matches_torch(hm):
    hm.hm_pow.restype = ctypes.c_double
    for g in (0.0, 0.3, 1.2, 1.5):
        for e in (2.0, 4.0, 6.0, 1.0, 2.5, 0.5):
            ours = hm.hm_pow(ctypes.c_double(g), ctypes.c_double(e))
            ref = torch.pow(torch.tensor(g, dtype=torch.float64), torch.tensor(e, dtype=torch.float64)).item()
            assert abs(ours - ref) <= 4e-16 * max(abs(ref), 1e-300)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/dc_b200.h declares."""
    header = open(os.path.join(ROOT, 'include', 'dc_b200.h')).read()
    declared = set(re.findall(r'^(?:int|const char\*)\s+(dc_\w+)\s*\(', header, flags=re.M))
    assert len(declared) >= 30
    from depth_correction_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared - {'dc_last_error', 'dc_version'} == set(_lib.SIGNATURES)
    assert _lib.version() >= 100
    # no torch types cross the boundary: every bound argument is a pointer, an integer or a double
    plain = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t,
             ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(_lib.GridSpec))
    for name, args in _lib.SIGNATURES.items():
        assert all(a in plain for a in args), name


def test_hot_path_refuses_cpu_tensors():
    import depth_correction_b200 as dc
    p = torch.rand(10, 3)
    with pytest.raises(RuntimeError):
        dc.nearest_neighbors(p, p, r=0.5)
    c = dc.DepthCloud.from_points(p)
    c.update_points()
    with pytest.raises(RuntimeError):
        c.update_neighbors(r=0.5)


def test_structured_array_round_trip_matches_reference():
    """DepthCloud <-> structured array (x, y, z, vp_*, normal_*, inc_angle, loss, mask; depth_cloud.py:508-533,
    577-590): field names, dtypes and values; live against the reference when its tree is present."""
    import depth_correction_b200 as dc
    from numpy.lib.recfunctions import unstructured_to_structured
    from oracle import ref_shim
    rng = np.random.default_rng(3)
    n = 50
    cols = np.concatenate([rng.normal(0, 5, (n, 3)), rng.normal(0, 0.1, (n, 3)), rng.normal(0, 1, (n, 3))], 1).astype(np.float32)
    cols[:, 6:9] /= np.linalg.norm(cols[:, 6:9], axis=1, keepdims=True)
    arr = unstructured_to_structured(cols, names=['x', 'y', 'z', 'vp_x', 'vp_y', 'vp_z', 'normal_x', 'normal_y', 'normal_z'])
    cloud = dc.DepthCloud.from_structured_array(arr)
    assert torch.allclose(cloud.to_points(), torch.as_tensor(cols[:, :3]), atol=1e-5)
    assert torch.equal(cloud.vps, torch.as_tensor(cols[:, 3:6])) and torch.equal(cloud.normals, torch.as_tensor(cols[:, 6:9]))
    cloud.inc_angles = torch.as_tensor(rng.uniform(0, 1.5, (n, 1)).astype(np.float32))
    cloud.loss = torch.as_tensor(rng.random(n).astype(np.float32))
    cloud.mask = torch.as_tensor(rng.random(n) < 0.5)
    out = cloud.to_structured_array()
    assert out.dtype.names == ('x', 'y', 'z', 'vp_x', 'vp_y', 'vp_z', 'normal_x', 'normal_y', 'normal_z', 'inc_angle', 'loss', 'mask')
    assert out['mask'].dtype == np.uint8 and out['x'].dtype == np.float32
    assert np.allclose(np.stack([out[f] for f in ('x', 'y', 'z')], 1), cols[:, :3], atol=1e-5)
    assert np.array_equal(out['inc_angle'], cloud.inc_angles.numpy()[:, 0]) and np.array_equal(out['mask'], cloud.mask.numpy().astype(np.uint8))
    if ref_shim.available():
        ref = ref_shim.load()
        rc = ref.DepthCloud.from_structured_array(arr)
        assert torch.allclose(rc.depth, cloud.depth) and torch.allclose(rc.dirs, cloud.dirs) and torch.equal(rc.vps, cloud.vps)
        rc.inc_angles = cloud.inc_angles
        rc.loss = cloud.loss[:, None]
        # (the reference's merge_arrays cannot take the uint8 mask field under numpy >= 2: compared without it)
        ro = rc.to_structured_array()
        assert ro.dtype.names == tuple(f for f in out.dtype.names if f != 'mask')
        for f in ro.dtype.names:
            assert ro[f].dtype == out[f].dtype and np.allclose(ro[f], out[f], atol=1e-6), f


def test_config_yaml_round_trip(tmp_path):
    """Config fields survive to_yaml / from_yaml (the training driver writes train.yaml / best.yaml with it)."""
    import depth_correction_b200 as dc
    cfg = dc.Config(nn_k=16, nn_r=0.3, eigenvalue_bounds=[[0, None, 0.01]], loss='trace_loss', lr=2e-3,
                    pose_correction=dc.PoseCorrection.pose, model_kwargs={'w': [0.0, 0.0], 'exponent': [2.0, 4.0]})
    path = str(tmp_path / 'cfg.yaml')
    cfg.to_yaml(path)
    back = dc.Config().from_yaml(path)
    for k in ('nn_k', 'nn_r', 'eigenvalue_bounds', 'loss', 'lr', 'pose_correction', 'model_kwargs', 'grid_res', 'random_seed'):
        assert getattr(back, k) == getattr(cfg, k), k
    copy = cfg.copy()
    copy.model_kwargs['w'][0] = 1.0                      # (one level deep, like the reference's Config.copy)
    assert isinstance(copy, dc.Config) and copy.nn_k == 16
