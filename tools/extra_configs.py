"""Measurements of the other BASELINE configs on one GPU (diagnosis + numbers for DESIGN.md; not the bench line).

  c1  latency of one S02-shaped graph (N=300, C=4, E=67,500): K0 + K1 + forward + decisions, p50/p99 over 300 calls
  c4  post-processing only on a large planted predicted graph (default 1M nodes), labels checked against the CPU oracle
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import gcn_mtmc_b200 as m

dev = torch.device("cuda", 0)


def c1(reps=300):
    net = bench.make_model(dev)
    x, ei = bench.device_graph(300, 4, 0, dev)
    b = bench.Batch(); b.x, b.edge_index, b.num_nodes = x, ei, 300

    def full():
        g = m.TrackletGraph(ei, 300)
        b._mpn_b200_graph = ((ei.data_ptr(), tuple(ei.shape), ei._version, 300, None), g)
        b.edge_attr = m.edge_features(x, ei, graph=g)
        net(b)

    def fwd_only():
        net(b)
    for name, fn in (("K0+K1+forward+decide", full), ("forward+decide only", fwd_only)):
        for _ in range(20):
            fn()
        ts = []
        for _ in range(reps):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record(); fn(); e.record(); e.synchronize()
            ts.append((a.elapsed_time(e), 1e3 * (time.perf_counter() - t0)))
        dv = sorted(t[0] for t in ts); hv = sorted(t[1] for t in ts)
        print("c1 %-22s device p50 %.3f ms p99 %.3f ms | host wall p50 %.3f ms  (E=%d)" %
              (name, dv[len(dv) // 2], dv[int(len(dv) * 0.99)], hv[len(hv) // 2], ei.shape[1]))


def c4(n_nodes=1_000_000, cams=8, extra=60.0, check=True):
    from oracle import postproc_oracle as po
    t0 = time.perf_counter()
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, 7, n_extra_per_node=extra, flip_on=0.02, flip_off=0.02, single_dir=0.01)
    print("c4 graph: N=%d E=%d active=%d (generated in %.1f s)" % (n_nodes, src.size, int(pred.sum()), time.perf_counter() - t0))
    ei = torch.from_numpy(np.stack([src, dst])).to(dev)
    d = bench.Batch(); d.x = torch.zeros(n_nodes, 1, device=dev); d.edge_index = ei; d.num_nodes = n_nodes
    p = torch.from_numpy(prob).to(dev)
    pr = torch.from_numpy(pred).to(dev)
    cfg = {"CUTTING": True, "PRUNING": True, "SPLITTING": True}
    m.graph_for(d, ei, n_nodes)
    for numbering in ("canonical", "reference"):
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ID, P = m.post_processing(cams, None, None, pr.clone(), None, dict(cfg), d, p, numbering=numbering)
            torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        print("c4 post_processing numbering=%-9s  best %.1f ms  -> %.2f G edges/s, clusters=%d, active after=%d" %
              (numbering, 1e3 * min(ts), src.size / min(ts) / 1e9, int(ID.max()) + 1, int(P.sum())))
    if check:
        t0 = time.perf_counter()
        if os.environ.get("MPN_C4_PYTHON_ORACLE") == "1":
            lab, act = po.post_processing_rounds(src, dst, pred, prob, cams, n_nodes, numbering="reference")
            print("c4 CPU oracle (numpy rounds): %.1f s" % (time.perf_counter() - t0))
        else:                                            # plain-C restatement: ~5x faster, pinned against the numpy one on the CPU
            from oracle import postproc_c as pc
            lab, act = pc.post_processing(src, dst, pred, prob, cams, n_nodes, numbering="reference")
            print("c4 CPU oracle (plain C, oracle/postproc_oracle.c): %.1f s" % (time.perf_counter() - t0))
        assert np.array_equal(P.cpu().numpy(), act), "decisions differ"
        assert np.array_equal(ID.numpy(), lab), "labels differ"
        print("c4 labels and decisions bit-exact vs oracle; max cluster size %d" % np.bincount(lab).max())


def c3(n_graphs=2000, n=300, cams=4):
    """BASELINE configs[2]: n_graphs S02-shaped graphs packed into one launch, BatchNorm statistics per graph."""
    net = bench.make_model(dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(n_graphs, n, bench.FEAT_DIM, generator=gen, device=dev)
    x = torch.nn.functional.normalize(x, p=2, dim=1).reshape(n_graphs * n, bench.FEAT_DIM)      # per graph, per column
    _, tmpl = bench.device_graph(n, cams, 0, dev)
    E1 = tmpl.shape[1]
    ei = (tmpl[None] + (torch.arange(n_graphs, device=dev) * n)[:, None, None]).permute(1, 0, 2).reshape(2, n_graphs * E1).contiguous()
    ptr = torch.arange(n_graphs + 1, device=dev) * n
    b = bench.Batch(); b.x, b.edge_index, b.num_nodes, b.ptr = x, ei, n_graphs * n, ptr

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = fn(); e.record(); e.synchronize()
            best = min(best, a.elapsed_time(e))
        return best, r
    t_g, g = timed(lambda: m.TrackletGraph(ei, n_graphs * n, ptr=ptr), 2)
    b._mpn_b200_graph = ((ei.data_ptr(), tuple(ei.shape), ei._version, n_graphs * n, (ptr.data_ptr(), ptr._version)), g)
    t_ef, ea = timed(lambda: m.edge_features(x, ei, graph=g))
    b.edge_attr = ea
    t_fw, (out, h) = timed(lambda: net(b))
    tot = t_g + t_ef + t_fw
    print("c3 %d graphs x (N=%d, E=%d): tables %.2f ms, edge features %.2f ms, forward+decide %.2f ms -> %.1f us/graph, %.2f G edges/s, %.0f graphs/s" %
          (n_graphs, n, E1, t_g, t_ef, t_fw, 1e3 * tot / n_graphs, ei.shape[1] / tot / 1e6, n_graphs / tot * 1e3))
    # parity spot check: graph k alone through the single-graph path
    worst = 0.0
    for k in (0, n_graphs // 3, n_graphs - 1):
        s = bench.Batch(); s.x = x[k * n:(k + 1) * n].contiguous(); s.edge_index = tmpl; s.num_nodes = n
        s.edge_attr = m.edge_features(s.x, tmpl)
        o1, h1 = net(s)
        d = (o1["classified_edges"][-1] - out["classified_edges"][-1][k * E1:(k + 1) * E1]).abs().max().item()
        worst = max(worst, d / o1["classified_edges"][-1].abs().max().item())
    print("c3 batched vs single-graph path: max |dlogit| / max|logit| = %.2e" % worst)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("c1", "all"):
        c1()
    if which in ("c3", "all"):
        c3(int(sys.argv[2]) if len(sys.argv) > 2 and which == "c3" else 2000)
    if which in ("c4", "all"):
        c4(int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000, extra=float(sys.argv[3]) if len(sys.argv) > 3 else 60.0)
