"""torch-profiler kernel table of the config-3 encoders (library modules) at the bench shape: python tools/profile_config3.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import hmvit_loader

pkg = hmvit_loader.load()
dev = torch.device("cuda:0")
enc = pkg.encoders
args = enc.config3_args()
torch.manual_seed(0)
cam = enc.CvtCameraEncoder(args['camera']).eval().to(dev)
pp = enc.PointPillar(args['lidar']).eval().to(dev).set_return_features()
n = 5
g = torch.Generator(device=dev).manual_seed(1)
batch = {'camera': torch.rand(n, 4, 512, 512, 3, device=dev, generator=g),
         'intrinsic': torch.tensor([[512., 0, 256], [0, 512., 256], [0, 0, 1]], device=dev).repeat(n, 4, 1, 1),
         'extrinsic': torch.eye(4, device=dev).repeat(n, 4, 1, 1)}
la = args['lidar']
nx, ny, _ = la['point_pillar_scatter']['grid_size']
M = 6000
coords = torch.cat([torch.stack([torch.full((M,), a, device=dev), torch.zeros(M, device=dev, dtype=torch.long),
                                 (c := torch.randperm(nx * ny, device=dev, generator=g)[:M]) // nx, c % nx], 1) for a in range(n)]).int()
vox = {'voxel_features': torch.rand(n * M, 32, 4, device=dev, generator=g), 'voxel_coords': coords,
       'voxel_num_points': torch.randint(1, 33, (n * M,), device=dev, generator=g).int()}
with torch.no_grad():
    for _ in range(2):
        cam(batch); pp({'processed_lidar': vox, 'batch_size': n})
    torch.cuda.synchronize()
    for name, fn in (("camera", lambda: cam(batch)), ("lidar", lambda: pp({'processed_lidar': vox, 'batch_size': n}))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        print(name, "ms", e0.elapsed_time(e1))
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn(); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
