"""Probe: how much of the configs[1] step is launch latency?  K1 + forward (graph tables prebuilt) timed eagerly and as one
CUDA-graph replay (torch.cuda.CUDAGraph around the same Python calls).  Diagnostic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import gcn_mtmc_b200 as m
dev = torch.device("cuda", 0)
net = bench.make_model(dev)
N = 4096
x, ei = bench.device_graph(N, 8, 0, dev)
g = m.TrackletGraph(ei, N)
b = bench.Batch(); b.num_nodes = N; b.x = x; b.mpn_graph = g
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def body():
    b.edge_attr = m.edge_features(x, None, graph=g)
    net(b)
    return net.last_pred

def timeit(fn, reps=20):
    ts = []
    for i in range(reps + 3):
        flush.fill_(i & 0xFF)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); e.synchronize()
        if i >= 3: ts.append(a.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2], ts[0]

print("eager  K1+forward: median %.3f ms  min %.3f ms" % timeit(body))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): body()
torch.cuda.current_stream().wait_stream(s)
cg = torch.cuda.CUDAGraph()
with torch.cuda.graph(cg):
    out = body()
ref = body().clone()
cg.replay(); torch.cuda.synchronize()
print("replay equals eager:", bool(torch.equal(out, ref)))
print("graph  K1+forward: median %.3f ms  min %.3f ms" % timeit(cg.replay))
