"""Device time of the tcgen05 apply sweep at configs[1] with 4 and 5 resident blocks per SM (MPN_ATC_CTAS is read per launch),
plus the whole bench step for each; measurement aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import gcn_mtmc_b200 as m

dev = torch.device("cuda", 0)
net = bench.make_model(dev)
N = 4096
x, ei = bench.device_graph(N, 8, 0, dev)
g = m.TrackletGraph(ei, N)
ea = m.edge_features(x, ei, graph=g)
W = net._weights(dev)
logits = torch.empty(1, g.n_edges, 2, device=dev)
pred = torch.empty(g.n_edges, dtype=torch.uint8, device=dev)
prob1 = torch.empty(g.n_edges, device=dev)
ph = m.CudaPhases(g, W, x, ea, 1, 1, g.n_edges, logits, pred, prob1, True, ws_kind="bench_plan")
S = m._lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
pre = [lambda: ph.node_encoder(), lambda: ph.sweep(0, S.STAGE_ENC0), lambda: ph.reduce(S.STAGE_ENC0, True),
       lambda: ph.sweep(0, S.STAGE_ENC1), lambda: ph.reduce(S.STAGE_ENC1, True), lambda: ph.node_tables(1),
       lambda: ph.sweep(1, S.STAGE_EDGE), lambda: ph.reduce(S.STAGE_EDGE, True), lambda: ph.sweep(1, S.STAGE_NODE),
       lambda: ph.reduce(S.STAGE_NODE, True)]
for f in pre:
    f()
ref = None
for ctas in ("4", "5", "4", "5"):
    os.environ["MPN_ATC_CTAS"] = ctas
    ts = []
    for rep in range(9):
        flush.fill_(rep)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ph.sweep(1, S.STAGE_APPLY, out_index=0, last=True)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ph.node_finalize(1)
    torch.cuda.synchronize()
    out = (logits.clone(), pred.clone(), ph.h_full().clone())
    if ref is None:
        ref = out
    same = all(torch.equal(a_, b_) for a_, b_ in zip(ref, out))
    ts = sorted(ts[1:])
    print("apply_tc ctas=%s  min %.4f ms  med %.4f ms  same_bits_as_first=%s" % (ctas, ts[0], ts[len(ts) // 2], same))
ph.close()
b_ = bench.Batch(); b_.num_nodes = N
for ctas in ("4", "5"):
    os.environ["MPN_ATC_CTAS"] = ctas
    ts = []
    for rep in range(8):
        flush.fill_(rep)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gg = m.TrackletGraph(ei, N, validate="deferred")
        b_.x, b_.mpn_graph, b_.edge_attr = x, gg, None
        net.fuse_decisions = True
        net(b_)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[1:])
    print("whole step ctas=%s  min %.4f ms  med %.4f ms" % (ctas, ts[0], ts[len(ts) // 2]))
