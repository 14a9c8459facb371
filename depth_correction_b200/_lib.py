"""ctypes binding of the C-ABI library declared in include/dc_b200.h (libdcb200.so, sm_100a).

There is deliberately no fallback: if the shared library is missing, importing this module
raises, and every hot-path entry point of the package fails loudly.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DC_B200_LIB') or os.path.join(_HERE, 'libdcb200.so')   # env override: kernel-variant experiments

DC_F32, DC_F64 = 0, 1
MODEL_NONE, MODEL_POLYNOMIAL, MODEL_SCALED_POLYNOMIAL = 0, 1, 2
MAX_TERMS = 8
LOSS_MIN_EIGVAL, LOSS_TRACE = 0, 1
FLAG_NORMALIZATION, FLAG_SQRT, FLAG_RAW = 1, 2, 4
SLICE = 32


class GridSpec(ctypes.Structure):
    _fields_ = [('origin', ctypes.c_double * 3), ('cell', ctypes.c_double),
                ('dims', ctypes.c_int32 * 3), ('axis', ctypes.c_int32 * 3), ('sub_bits', ctypes.c_int32)]


class DcError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        'depth_correction_b200: %s not found. Build it with `python -c "import __graft_entry__ as g; g.build()"` '
        'or `make -C depth_correction_b200/csrc`. There is no CPU or PyTorch fallback for the hot path.' % LIB_PATH)

_lib = ctypes.CDLL(LIB_PATH)
_lib.dc_last_error.restype = ctypes.c_char_p
_lib.dc_version.restype = ctypes.c_int

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_D = ctypes.c_double
_SZ = ctypes.c_size_t
_SZP = ctypes.POINTER(ctypes.c_size_t)
_SPEC = ctypes.POINTER(GridSpec)

# name -> argtypes, exactly the declarations of include/dc_b200.h
SIGNATURES = {
    'dc_bounds': [_P, _I, _L, _P, _P, _P],
    'dc_cell_keys': [_P, _I, _L, _SPEC, _P, _P, _P],
    'dc_cell_keys_stacked': [_P, _I, _L, _SPEC, _P, _I, _I, _I, _P, _P, _P],
    'dc_sort_pairs': [_P, _P, _P, _P, _L, _I, _P, _SZP, _P],
    'dc_gather_points': [_P, _I, _P, _L, _P, _P, _P],
    'dc_cell_table': [_P, _L, _L, _I, _P, _P],
    'dc_radius_count': [_P, _P, _L, _P, _P, _L, _SPEC, _P, _D, _P, _P, _P],
    'dc_ell_offsets': [_P, _L, _P, _P, _SZP, _P],
    'dc_radius_fill': [_P, _P, _L, _P, _P, _L, _SPEC, _P, _D, _P, _P, _P],
    'dc_knn': [_P, _P, _L, _P, _P, _L, _SPEC, _P, _I, _D, _P, _P, _P],
    'dc_knn_cells': [_P, _P, _L, _P, _P, _L, _SPEC, _P, _I, _D, _P, _P, _SZP, _P],
    'dc_knn_recorded': [_P, _P, _L, _P, _P, _L, _SPEC, _P, _I, _D, _P, _P, _SZP, _P],
    'dc_knn_sort_rows': [_P, _L, _I, _P, _P, _L, _P],
    'dc_knn_distances': [_P, _P, _I, _P, _L, _P, _P],
    'dc_ell_to_padded': [_P, _P, _L, _P, _P, _I, _P, _P],
    'dc_ell_to_dist': [_I, _P, _P, _L, _P, _P, _P],
    'dc_sort_rows': [_P, _L, _I, _P, _SZP, _P],
    'dc_padded_to_ell': [_P, _L, _I, _P, _P, _P, _P, _P, _P],
    'dc_graph_edges': [_P, _P, _L, _P, _P, _P],
    'dc_graph_degrees': [_P, _P, _L, _P, _P],
    'dc_sort_keys': [_P, _P, _L, _I, _I, _P, _SZP, _P],
    'dc_exclusive_sum_i32_i64': [_P, _P, _L, _P, _SZP, _P],
    'dc_transpose_widths': [_P, _L, _L, _P, _P, _P],
    'dc_transpose_fill': [_P, _L, _L, _P, _P, _P],
    'dc_pack_records': [_P, _P, _P, _P, _P, _P, _I, _L, _L, _I, _P, _P, _P, _P, _P],
    'dc_pack_records_batched': [_P, _P, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    'dc_world_points_batched': [_P, _P, _I, _L, _I, _P, _P, _P],
    'dc_set_loss_mask': [_P, _L, _P, _P, _P],
    'dc_step_points': [_P, _P, _P, _I, _L, _P, _I, _I, _P, _P, _I, _P, _P],
    'dc_step_forward': [_P, _P, _L, _P, _P, _I, _I, _P, _P, _P, _P, _P, _SZ, _P],
    'dc_step_forward_scatter': [_P, _P, _L, _P, _P, _I, _I, _P, _P, _P, _P, _SZ, _P],
    'dc_step_backward': [_P, _L, _P, _P, _P, _P, _P, _P, _P],
    'dc_step_backward_scatter': [_P, _L, _P, _P, _P, _P, _P, _I, _P],
    'dc_step_chain': [_P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P],
    'dc_pose_compose': [_P, _P, _I, _I, _P, _P],
    'dc_pose_compose_backward': [_P, _P, _I, _I, _P, _P, _P],
    'dc_features': [_P, _I, _L, _P, _P, _I, _P, _P, _P],
    'dc_local_features_finish': [_P, _P, _P, _P, _I, _L, _I, _P, _P, _P, _P, _P],
    'dc_feature_mask': [_P, _I, _L, _I, _P, _I, _P, _L, _I, _P, _P],
    'dc_features_backward': [_P, _I, _L, _P, _P, _I, _P, _P, _P, _P],
    'dc_eigh3': [_P, _I, _L, _P, _P, _P],
    'dc_eigh3_backward': [_P, _P, _I, _L, _P, _P, _P, _P],
    'dc_normals_angles': [_P, _P, _I, _L, _I, _P, _P, _P],
    'dc_world_points': [_P, _P, _P, _I, _L, _P, _P, _P],
    'dc_from_points': [_P, _P, _I, _L, _P, _P, _P, _P],
    'dc_voxel_keys': [_P, _I, _L, _D, _P, _I, _P, _P, _P, _P],
    'dc_voxel_pick': [_P, _P, _L, _P, _I, _I, _P, _P, _P, _P],
    'dc_icp_forward': [_P, _P, _I, _P, _P, _I, _P, _P, _D, _P, _P, _L, _I, _P, _P, _SZ, _P],
    'dc_icp_backward': [_P, _P, _I, _P, _P, _I, _P, _P, _D, _P, _P, _L, _I, _P, _P, _P, _P, _P, _P],
    'dc_f64_sort_keys': [_P, _L, _P, _P, _P],
    'dc_f64_from_sort_keys': [_P, _L, _P, _P],
    'dc_route_count': [_P, _I, _L, _P, _I, _D, _P, _P, _I, _P, _P, _P, _P, _I, _P],
    'dc_route_pack': [_P, _P, _P, _I, _L, _I, _P, _I, _P, _I, _D, _P, _P, _P, _P, _P, _P, _P],
    'dc_axis_histogram': [_P, _I, _L, _D, _D, _I, _P, _P],
    'dc_route_keys': [_P, _L, _P, _P, _P],
    'dc_route_unpack': [_P, _P, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    'dc_neighbor_stats': [_P, _P, _I, _P, _P, _L, _I, _P, _P, _P],
    'dc_shadow_mask': [_P, _P, _I, _P, _P, _L, _I, _D, _D, _P, _P, _P, _P],
}

for _name, _args in SIGNATURES.items():
    _f = getattr(_lib, _name)   # AttributeError here == the library does not export a declared symbol
    _f.argtypes = _args
    _f.restype = ctypes.c_int

# number of kernel launches issued through this binding (bench.py reports it as gpu_launches)
launch_count = 0


def version():
    return _lib.dc_version()


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, 'depth_correction_b200 kernels need CUDA tensors (no CPU fallback)'
    assert t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def upload(host, dtype, device):
    """Small host table -> device without draining the stream: a pageable cudaMemcpy waits for everything already
    queued, a copy from pinned memory (torch's caching host allocator recycles the staging block in stream order) is
    just another stream operation."""
    t = torch.as_tensor(host, dtype=dtype)
    if torch.device(device).type != 'cuda':
        return t
    return t.pin_memory().to(device, non_blocking=True)


def dtype_code(dt):
    if dt == torch.float32:
        return DC_F32
    if dt == torch.float64:
        return DC_F64
    raise TypeError('depth_correction_b200 supports float32 / float64 clouds, got %s' % dt)


# optional per-entry-point CUDA-event timing (bench.py): set `profile = {}` to enable, None to disable
profile = None


def call(name, *args):
    global launch_count
    if profile is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(_lib, name)(*args)
    if rc != 0:
        raise DcError('%s failed (code %d): %s' % (name, rc, _lib.dc_last_error().decode()))
    launch_count += 1
    if profile is not None:
        e1.record()
        profile.setdefault(name, []).append((e0, e1))


def collect_profile():
    """{entry point: {'calls': n, 'ms_total': t}} from the recorded events (call after a synchronize)."""
    out = {}
    for name, evs in (profile or {}).items():
        out[name] = {'calls': len(evs), 'ms_total': float(sum(a.elapsed_time(b) for a, b in evs))}
    return out


# Reusable scratch buffers, one per (tag, device), grown to the largest request.  Reuse is stream-ordered
# (everything runs on the current stream), so a scratch buffer may be handed out again as soon as the
# kernels that used it have been enqueued.  Keeps multi-GB temporaries (radix-sort double buffers, edge
# pairs) out of the caching allocator's split/merge churn when the search is repeated.
_workspace = {}


def scratch(tag, numel, dtype, device):
    nbytes = int(numel) * torch.empty((), dtype=dtype).element_size()
    key = (tag, str(device))
    buf = _workspace.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspace.pop(key, None)
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _workspace[key] = buf
    return buf[:nbytes].view(dtype)


def release_workspace():
    _workspace.clear()


def call_with_temp(name, device, *args_before_temp, after=()):
    """Two-phase (temp, temp_bytes) protocol: size query with temp == NULL, then the real call."""
    nbytes = ctypes.c_size_t(0)
    global launch_count
    rc = getattr(_lib, name)(*args_before_temp, None, ctypes.byref(nbytes), *after)
    if rc != 0:
        raise DcError('%s (size query) failed (code %d): %s' % (name, rc, _lib.dc_last_error().decode()))
    temp = scratch('temp:' + name, max(int(nbytes.value), 1), torch.uint8, device)
    nbytes = ctypes.c_size_t(temp.numel())
    call(name, *args_before_temp, ptr(temp), ctypes.byref(nbytes), *after)
    return temp
