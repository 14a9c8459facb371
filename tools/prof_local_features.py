import sys, time, torch
sys.path.insert(0, '.')
import depth_correction_b200 as dc
from bench import host_scans, local_features, NN_K, NN_R
dev = torch.device('cuda:0')
pts_host, poses_np = host_scans(64, 'os0-128')
cfg = dc.Config(nn_k=NN_K, nn_r=NN_R, pose_correction=dc.PoseCorrection.pose)
pts_dev = [torch.from_numpy(p).to(dev) for p in pts_host]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    clouds = local_features(dc, pts_dev, cfg)
    torch.cuda.synchronize(); print('local_feature_cloud x64 (kNN k=32 r=0.4 per scan): %.1f ms' % ((time.perf_counter() - t0) * 1e3))
cfg2 = dc.Config(nn_k=0, nn_r=0.25, grid_res=0.1, min_depth=1.0, max_depth=25.0)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cl = [dc.local_feature_cloud(dc.filtered_cloud(dc.DepthCloud.from_points(p), cfg2), cfg2) for p in pts_dev]
    torch.cuda.synchronize(); print('filtered_cloud + local_feature_cloud x64 (grid 0.1, r=0.25): %.1f ms, %d points' % ((time.perf_counter() - t0) * 1e3, sum(len(c) for c in cl)))
