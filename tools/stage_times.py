"""Host wall-clock (with device sync) of each stage of one bench step; diagnosis aid, not a benchmark."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import gcn_mtmc_b200 as m

dev = torch.device("cuda", 0)
net = bench.make_model(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x, ei = bench.device_graph(N, 8, 0, dev)
b = bench.Batch(); b.x, b.edge_index, b.num_nodes = x, ei, N


def timed(name, fn, reps=5):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("%-22s min %.3f ms  med %.3f ms  max %.3f ms" % (name, 1e3 * min(ts), 1e3 * sorted(ts)[len(ts) // 2], 1e3 * max(ts)))
    return r


timed("require_device", lambda: m._lib.require_device(0))
g = timed("TrackletGraph", lambda: m.TrackletGraph(ei, N))
ea = timed("edge_features", lambda: m.edge_features(x, ei, graph=g))
b.edge_attr = ea
b._mpn_b200_graph = ((ei.data_ptr(), tuple(ei.shape), ei._version, N, None), g)
timed("forward", lambda: net(b))
timed("weights()", lambda: net._weights(dev))
