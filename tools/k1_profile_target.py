"""Target for ncu: two calls of the fused edge-feature path at configs[1] size (tables from camera ids)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import gcn_mtmc_b200 as m
dev = torch.device("cuda", 0)
N, cams = bench.NODES_1GPU, bench.CAMS
x, ei = bench.device_graph(N, cams, 0, dev)
cam_host = (torch.arange(N) * cams // N).numpy()
g = m.TrackletGraph.from_cameras(cam_host, dev)
for _ in range(2):
    ea = m.edge_features(x, None, graph=g)
torch.cuda.synchronize()
print("ok", float(ea.sum()))
