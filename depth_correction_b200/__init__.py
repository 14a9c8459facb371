"""depth_correction_b200 -- B200-native (sm_100a) implementation of the map-consistency training
hot path of ctu-vras/depth_correction behind the reference's Python API.

    from depth_correction_b200 import DepthCloud, ScaledPolynomial, min_eigval_loss, ...

Importing the package loads libdcb200.so (the C-ABI library of include/dc_b200.h); if it has not
been built the import fails -- there is no CPU or PyTorch fallback for the hot path.
"""
from . import _lib                                    # noqa: F401  (fails loudly when the library is missing)
from .config import Config, Loss, Model, NeighborhoodType, PoseCorrection
from .depth_cloud import DepthCloud
from .nearest_neighbors import ball_angle_to_distance, nearest_neighbors
from .model import BaseModel, Polynomial, ScaledPolynomial, load_model, model_by_name
from .loss import Reduction, batch_loss, create_loss, fused_sum_count, loss_by_name, min_eigval_loss, reduce, trace_loss
from .icp import icp_loss, point_to_plane_dist, point_to_point_dist
from .parallel import LocalMap, SlabPartitioner, distributed_quantile, reduce_step, sharded_inlier_sum_count
from .filters import (feature_mask, filter_depth, filter_eigenvalue, filter_eigenvalue_ratio, filter_eigenvalue_ratios,
                      filter_eigenvalues, filter_shadow_points, filter_valid_neighbors, within_bounds)
from .filters_grid import filter_grid
from .fused import set_backward_form
from .capture import CapturedIteration
from .preproc import (GlobalCloud, Neighborhoods, compute_neighborhood_features, establish_neighborhoods,
                      filtered_cloud, global_cloud, global_cloud_mask, local_feature_cloud, local_feature_clouds, offset_cloud)
from .eval import create_corrected_poses, eval_loss_clouds, initialize_pose_corrections
from .train import TrainCallbacks, train
from .transform import matrix_to_xyz_axis_angle, xyz_axis_angle_to_matrix
from .utils import covs, trace

__version__ = '0.1.0'
