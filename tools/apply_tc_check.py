"""Diagnostics for the apply sweep: the same forwards with the tensor-core kernel (MPN_APPLY_TC=1) and the packed-fp32 kernel
(MPN_APPLY_TC=0), on stored y, in separate processes (the switches are read once); compares h / logits and prints the
per-phase times of configs[1].  Not a benchmark."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = "/tmp"


def child(mode):
    import numpy as np, torch
    import bench
    import gcn_mtmc_b200 as m
    dev = torch.device("cuda", 0)
    res = {}
    for tag, N, L, ncls, chunk in (("a", 2048, 1, 1, None), ("b", 1024, 3, 2, 128), ("c", 600, 2, 1, 256)):
        net = bench.make_model(dev, L, ncls)
        x, ei = bench.device_graph(N, 8 if tag != "c" else 3, 0, dev)
        g = m.TrackletGraph(ei, N, chunk=chunk)
        b = bench.Batch(); b.x, b.edge_index, b.num_nodes = x, ei, N
        b.edge_attr = m.edge_features(x, ei, graph=g)
        b.mpn_graph = g
        out, h = net(b)
        torch.cuda.synchronize()
        res[tag + "_h"] = h.cpu().numpy()
        res[tag + "_lg"] = out["classified_edges"][-1].cpu().numpy()
        print(mode, tag, "chunk", g.chunk, "E", g.n_edges, "h absmax", float(h.abs().max()), flush=True)
    np.savez(os.path.join(OUT, "atc_%s.npz" % mode), **res)
    net = bench.make_model(dev)
    x, ei = bench.device_graph(4096, 8, 0, dev)
    ph = bench.time_phases(m, net, x, ei)
    print(mode, "phases", {k: round(v, 4) for k, v in ph.items()}, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
        sys.exit(0)
    import numpy as np
    for mode in ("0", "1"):
        r = subprocess.run([sys.executable, __file__, mode], env=dict(os.environ, MPN_APPLY_TC=mode, MPN_STORE_Y="1"), timeout=600)
        print("MPN_APPLY_TC=%s rc %d" % (mode, r.returncode), flush=True)
    ref, z = np.load(os.path.join(OUT, "atc_0.npz")), np.load(os.path.join(OUT, "atc_1.npz"))
    for k in ref.files:
        a, b = ref[k].astype(np.float64), z[k].astype(np.float64)
        print("%-5s max|tensor-core - packed fp32| %.3e  (max|value| %.3e)" % (k, np.abs(a - b).max(), np.abs(a).max()))
