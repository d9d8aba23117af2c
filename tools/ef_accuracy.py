"""Accuracy of the edge features at scale: our kernel (3xFP16 or 3xTF32 Gram planes, MPN_GRAM_F16=1/0) against an fp64
evaluation of inference.py:453-456 on the device (torch, chunked).  Diagnostic; prints max / rms errors."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    import bench
    import gcn_mtmc_b200 as m
    from oracle import mpn_oracle as mo
    dev = torch.device("cuda", 0)
    for tag, N, planted in (("random", 2048, False), ("planted", 2048, True), ("random4k", 4096, False)):
        if planted:
            x, ei, _, _ = mo.synth_graph(N, 8, 1, planted=True)
            x, ei = x.to(dev), ei.to(dev)
        else:
            x, ei = bench.device_graph(N, 8, 0, dev)
        g = m.TrackletGraph(ei, N)
        ea = m.edge_features(x, ei, graph=g).double()
        xd = x.double()
        E = ei.shape[1]
        worst = [0.0, 0.0]
        sq = [0.0, 0.0]
        for s in range(0, E, 1 << 18):
            a, b = xd[ei[0, s:s + (1 << 18)]], xd[ei[1, s:s + (1 << 18)]]
            d = (a - b + 1e-6).norm(dim=1)
            c = 1 - (a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1)).clamp_min(1e-8)
            e0, e1 = (ea[s:s + (1 << 18), 0] - d).abs(), (ea[s:s + (1 << 18), 1] - c).abs()
            worst = [max(worst[0], e0.max().item()), max(worst[1], e1.max().item())]
            sq = [sq[0] + (e0 ** 2).sum().item(), sq[1] + (e1 ** 2).sum().item()]
        std = ea.std(dim=0)
        print("%s F16=%s N=%d E=%d  dist: max %.2e rms %.2e (std of feature %.3e)   1-cos: max %.2e rms %.2e (std %.3e)" %
              (tag, os.environ.get("MPN_GRAM_F16", "1"), N, E, worst[0], (sq[0] / E) ** 0.5, std[0].item(), worst[1], (sq[1] / E) ** 0.5,
               std[1].item()), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for mode in ("1", "0"):
            subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MPN_GRAM_F16=mode), timeout=600)
