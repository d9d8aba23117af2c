"""One configs[1] step (K0 + K1 + forward + decisions) between cudaProfilerStart/Stop after warm-up, for
`ncu --profile-from-start off`.  Every kernel of the step appears exactly once in the capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import gcn_mtmc_b200 as m
dev = torch.device("cuda", 0)
net = bench.make_model(dev)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x, ei = bench.device_graph(N, 8, 0, dev)
b = bench.Batch(); b.num_nodes = N


def step():                      # the bench step: tables from the int64 edge_index, ONE forward call with the edge features inside
    g = m.TrackletGraph(ei, N, validate="deferred")
    b.x, b.edge_index, b.mpn_graph, b.edge_attr = x, ei, g, None
    net(b)
    g.validate()
    return net.last_pred


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
p = step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("active edges", int(p.sum().item()))
