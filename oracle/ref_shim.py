"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference under /root/reference.

The reference (ctu-vras/depth_correction) is pure Python but imports ROS, open3d,
matplotlib and pytorch3d at module top level (depth_cloud.py:4-11, utils.py:7-8,
transform.py:2-10, loss.py:14).  None of those touch the map-consistency path, so
empty stub modules are registered for them and the reference modules are then
imported unmodified from /root/reference/src.  Only `oracle/make_golden.py` and the
in-container cross-check tests use this file; /root/reference does not exist on the
GPU box, so nothing in `-m gpu` tests, `smoke()` or `bench.py` may import it.

pytorch3d is not installed: `axis_angle_to_matrix` is supplied from
`oracle.oracle.axis_angle_to_matrix` (a restatement of pytorch3d's published
axis-angle -> quaternion -> matrix formula; pinned against scipy's
Rotation.from_rotvec in tests/test_oracle.py).  That single function is therefore
"parity unpinned" against pytorch3d itself.
"""
import os
import sys
import types

import numpy as np

REFERENCE_SRC = '/root/reference/src'


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'depth_correction'))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Dummy(object):
    def __init__(self, *a, **k):
        pass


def install_stubs():
    from . import oracle as _oracle

    if not hasattr(np, 'object'):
        np.object = object  # nearest_neighbors.py:69 uses the removed alias
    if 'matplotlib' not in sys.modules:
        mpl = _stub('matplotlib')
        mpl.cm = _stub('matplotlib.cm', gist_rainbow=None, viridis=None)
        mpl.colors = _stub('matplotlib.colors', BASE_COLORS={})
        mpl.pyplot = _stub('matplotlib.pyplot')
    _stub('ros_numpy', msgify=None, numpify=None)
    _stub('rospy')
    sm = _stub('sensor_msgs')
    sm.msg = _stub('sensor_msgs.msg', PointCloud2=_Dummy)
    gm = _stub('geometry_msgs')
    gm.msg = _stub('geometry_msgs.msg', Point=_Dummy, Pose=_Dummy, PoseStamped=_Dummy, Quaternion=_Dummy,
                   Transform=_Dummy, TransformStamped=_Dummy)
    nm = _stub('nav_msgs')
    nm.msg = _stub('nav_msgs.msg', Path=_Dummy)
    st = _stub('std_msgs')
    st.msg = _stub('std_msgs.msg', Header=_Dummy)
    _stub('open3d')
    p3 = _stub('pytorch3d')
    p3.io = _stub('pytorch3d.io', load_ply=None, load_obj=None, IO=_Dummy)
    p3.structures = _stub('pytorch3d.structures', Meshes=_Dummy, Pointclouds=_Dummy)
    p3.ops = _stub('pytorch3d.ops')
    p3.ops.knn = _stub('pytorch3d.ops.knn', knn_points=knn_points_exact)
    p3.transforms = _stub('pytorch3d.transforms',
                          axis_angle_to_matrix=_oracle.axis_angle_to_matrix,
                          matrix_to_quaternion=None,
                          quaternion_to_axis_angle=None,
                          axis_angle_to_quaternion=_oracle.axis_angle_to_quaternion)


def knn_points_exact(p1, p2, K=1):
    """Stand-in for pytorch3d.ops.knn.knn_points (not vendored, not installed; the reference's ICP losses call it
    at loss.py:445, 532 with K=1): batched nearest neighbours of p1 [1,n1,3] in p2 [1,n2,3] -> (squared distances
    [1,n1,K], indices [1,n1,K], None).  pytorch3d searches by brute force in the tensors' dtype; this stand-in
    searches exactly (scipy cKDTree on the same values), so it differs from pytorch3d only for float32 near-ties
    ("parity unpinned" against pytorch3d itself).  Distances stay differentiable like pytorch3d's."""
    import numpy as np
    import torch
    from scipy.spatial import cKDTree
    assert K == 1 and p1.dim() == 3 and p1.shape[0] == 1
    a = p1[0].detach().cpu().numpy().astype(np.float64)
    b = p2[0].detach().cpu().numpy().astype(np.float64)
    _, idx = cKDTree(b).query(a, k=1)
    idx = torch.as_tensor(idx, dtype=torch.int64)
    d2 = ((p1[0] - p2[0][idx]) ** 2).sum(dim=-1)
    return d2[None, :, None], idx[None, :, None], None


def load():
    """Import the reference package and return a namespace with the hot-path symbols."""
    assert available(), 'reference tree not present (expected on the GPU box)'
    install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import depth_correction.config as ref_config
    ref_config.cmd_out = lambda *a, **k: ('', '')  # Config() shells out to git (config.py:160)
    from depth_correction.depth_cloud import DepthCloud
    from depth_correction.nearest_neighbors import nearest_neighbors
    from depth_correction.utils import covs, trace
    from depth_correction.model import Polynomial, ScaledPolynomial
    from depth_correction.loss import min_eigval_loss, trace_loss, Reduction
    from depth_correction import filters
    from depth_correction.transform import xyz_axis_angle_to_matrix
    import depth_correction.preproc as preproc
    ns = types.SimpleNamespace(
        DepthCloud=DepthCloud, nearest_neighbors=nearest_neighbors, covs=covs, trace=trace,
        Polynomial=Polynomial, ScaledPolynomial=ScaledPolynomial,
        min_eigval_loss=min_eigval_loss, trace_loss=trace_loss, Reduction=Reduction,
        filters=filters, xyz_axis_angle_to_matrix=xyz_axis_angle_to_matrix,
        global_cloud=preproc.global_cloud,
        compute_neighborhood_features=preproc.compute_neighborhood_features,
        establish_neighborhoods=preproc.establish_neighborhoods,
        local_feature_cloud=preproc.local_feature_cloud,
        global_cloud_mask=preproc.global_cloud_mask,
        NeighborhoodType=ref_config.NeighborhoodType,
        PoseCorrection=ref_config.PoseCorrection,
        config=ref_config,
    )
    return ns
