"""Target for ncu: three one-call steps at configs[1] size (tables from camera ids -> MOTMPNet.forward with edge_attr=None)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import gcn_mtmc_b200 as m
dev = torch.device("cuda", 0)
N, cams = bench.NODES_1GPU, bench.CAMS
x, ei = bench.device_graph(N, cams, 0, dev)
cam_host = (torch.arange(N) * cams // N).numpy()
net = bench.make_model(dev)
for _ in range(3):
    b = bench.Batch()
    b.x, b.num_nodes, b.edge_attr = x, N, None
    b.mpn_graph = m.TrackletGraph.from_cameras(cam_host, dev)
    net(b)
torch.cuda.synchronize()
print("ok", int(net.last_pred.sum()))
