"""Round-2 diagnostic of the fused edge-feature kernel (csrc/gram_ef.cu) at BASELINE configs[1] size: accuracy against an fp64
evaluation of inference.py:453-456 on every edge, device time of K1 alone (camera-built and edge_index-built graph tables) and of
the one-call step (tables from camera ids -> MOTMPNet.forward with edge_attr=None), L2 flushed between repetitions.

    python tools/k1_check.py [N [cams]]          # writes gpurun_out/k1_check.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gcn_mtmc_b200 as m


def timeit(fn, flush, reps=20):
    ts = []
    for i in range(reps + 3):
        flush.fill_(i & 0xFF)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        e.record()
        e.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else bench.NODES_1GPU
    cams = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CAMS
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m._lib.require_device(0)
    x, ei = bench.device_graph(N, cams, 0, dev)
    E = ei.shape[1]
    cam_host = (torch.arange(N) * cams // N).numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"N": N, "cams": cams, "E": E}
    g_cam = m.TrackletGraph.from_cameras(cam_host, dev)
    g_ei = m.TrackletGraph(ei, N)
    ea = m.edge_features(x, None, graph=g_cam)
    ea2 = m.edge_features(x, None, graph=g_ei)
    torch.cuda.synchronize()
    out["edge_index_graph_bit_identical"] = bool(torch.equal(ea, ea2))
    # accuracy on every edge against fp64
    xd = x.double()
    worst, sq = [0.0, 0.0], [0.0, 0.0]
    for s in range(0, E, 1 << 18):
        a, b = xd[ei[0, s:s + (1 << 18)]], xd[ei[1, s:s + (1 << 18)]]
        d = (a - b + 1e-6).norm(dim=1)
        c = 1 - (a * b).sum(1) / (a.norm(dim=1) * b.norm(dim=1)).clamp_min(1e-8)
        e0, e1 = (ea[s:s + (1 << 18), 0].double() - d).abs(), (ea[s:s + (1 << 18), 1].double() - c).abs()
        worst = [max(worst[0], e0.max().item()), max(worst[1], e1.max().item())]
        sq = [sq[0] + (e0 ** 2).sum().item(), sq[1] + (e1 ** 2).sum().item()]
    out["fp64_check"] = {"dist_max_abs": worst[0], "dist_rms": (sq[0] / E) ** 0.5, "cos_max_abs": worst[1], "cos_rms": (sq[1] / E) ** 0.5}
    print(json.dumps(out), flush=True)
    out["k1_ms_cameras"] = timeit(lambda: m.edge_features(x, None, graph=g_cam), flush)
    out["k1_ms_edge_index_tables"] = timeit(lambda: m.edge_features(x, None, graph=g_ei), flush)
    out["k1_ms_simt_gather"] = timeit(lambda: m.edge_features(x, None, graph=g_cam, use_tensor_cores=False), flush, reps=3)
    net = bench.make_model(dev)

    def step_cameras():
        b = bench.Batch()
        b.x, b.num_nodes, b.edge_attr = x, N, None
        b.mpn_graph = m.TrackletGraph.from_cameras(cam_host, dev)
        net(b)
        return net.last_pred

    def step_edge_index():
        b = bench.Batch()
        b.x, b.edge_index, b.num_nodes, b.edge_attr = x, ei, N, None
        b.mpn_graph = g = m.TrackletGraph(ei, N, validate="deferred")
        net(b)
        g.validate()
        return net.last_pred

    def step_two_calls():
        b = bench.Batch()
        b.x, b.edge_index, b.num_nodes = x, ei, N
        b.mpn_graph = g = m.TrackletGraph(ei, N, validate="deferred")
        b.edge_attr = m.edge_features(x, None, graph=g)
        net(b)
        g.validate()
        return net.last_pred

    p0, p1, p2 = step_cameras().clone(), step_edge_index().clone(), step_two_calls().clone()
    torch.cuda.synchronize()
    out["decisions_differ_cameras_vs_edge_index"] = int((p0 != p1).sum().item())
    out["decisions_differ_fused_moments_vs_sweep"] = int((p0 != p2).sum().item())
    out["step_ms_cameras"] = timeit(step_cameras, flush)
    out["step_ms_edge_index"] = timeit(step_edge_index, flush)
    out["step_ms_two_calls"] = timeit(step_two_calls, flush)
    lib = m._lib.lib()
    lib.mpn_set_pdl(0)
    out["step_ms_cameras_no_pdl"] = timeit(step_cameras, flush)
    lib.mpn_set_pdl(1)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/k1_check.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
