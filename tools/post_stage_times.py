"""Per-stage wall time and round counts of the GPU post-processing (diagnosis)."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import gcn_mtmc_b200 as m
from gcn_mtmc_b200.postprocess import _Post
from oracle import postproc_oracle as po

dev = torch.device("cuda", 0)
n_nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, 8, 7, n_extra_per_node=float(sys.argv[2]) if len(sys.argv) > 2 else 60.0)
ei = torch.from_numpy(np.stack([src, dst])).to(dev)
d = bench.Batch(); d.x = torch.zeros(n_nodes, 1, device=dev); d.edge_index = ei; d.num_nodes = n_nodes
p = torch.from_numpy(prob).to(dev); pr = torch.from_numpy(pred).to(dev)
m.graph_for(d, ei, n_nodes)
L = m._lib.lib()


def T(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("%-28s %.2f ms %s" % (name, 1e3 * (time.perf_counter() - t0), r if r is not None else ""))


for rep in range(2):
    P = _Post(d, pr.clone(), p)
    T("cut#1 (incl. list build)", lambda: m._lib.check(L.mpn_cut(P.g.ref, P.act.data_ptr(), P.ws.data_ptr(), P.ws.numel(), P.stream)))
    ch, rd = C.c_int32(0), C.c_int32(0)
    T("prune", lambda: (m._lib.check(L.mpn_prune(P.g.ref, P.act.data_ptr(), P.prob_ptr, P.prob_stride, 8, C.byref(ch), C.byref(rd), P.ws.data_ptr(), P.ws.numel(), P.stream)), "rounds=%d" % rd.value)[1])
    T("cut#2", lambda: m._lib.check(L.mpn_cut(P.g.ref, P.act.data_ptr(), P.ws.data_ptr(), P.ws.numel(), P.stream)))
    T("split", lambda: (m._lib.check(L.mpn_split(P.g.ref, P.act.data_ptr(), P.prob_ptr, P.prob_stride, 8, C.byref(rd), P.ws.data_ptr(), P.ws.numel(), P.stream)), "rounds=%d" % rd.value)[1])
    T("scc labels canonical", lambda: P.labels_canonical()[1])
    T("labels reference (host)", lambda: P.labels_reference()[1])
    print("active now", int(P.act.sum()))
