"""Where the configs[1] step's time goes beyond its kernels: the one-call step (tables from camera ids -> forward with
edge_attr=None -> decisions) launched eagerly, replayed as one CUDA graph (no host launch cost, same kernels), and the e2e
GraphStream with and without replay.  Run once per setting of MPN_GRAM_BALANCE (read once per process).

    MPN_GRAM_BALANCE=0 python tools/step_gaps.py ; MPN_GRAM_BALANCE=1 python tools/step_gaps.py     # gpurun_out/step_gaps_<b>.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gcn_mtmc_b200 as m
from k1_check import timeit


def main():
    N, cams = bench.NODES_1GPU, bench.CAMS
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m._lib.require_device(0)
    x, ei = bench.device_graph(N, cams, 0, dev)
    cam_host = (torch.arange(N) * cams // N).numpy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    net = bench.make_model(dev)
    out = {"balance": os.environ.get("MPN_GRAM_BALANCE", "1"), "E": int(ei.shape[1])}

    def step_cameras():
        b = bench.Batch()
        b.x, b.num_nodes, b.edge_attr = x, N, None
        b.mpn_graph = m.TrackletGraph.from_cameras(cam_host, dev)
        net(b)

    out["step_ms_cameras"] = timeit(step_cameras, flush)
    lib = m._lib.lib()
    lib.mpn_profile_gram(1)
    step_cameras()
    torch.cuda.synchronize()
    out["gram_ms_in_step"] = lib.mpn_profile_gram_ms()
    lib.mpn_profile_gram(0)
    g = m.TrackletGraph.from_cameras(cam_host, dev)
    out["k1_ms_alone"] = timeit(lambda: m.edge_features(x, None, graph=g), flush)
    # back-to-back throughput (no flush): host enqueue rate vs device rate
    for reps in (50,):
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            step_cameras()
        e.record()
        e.synchronize()
        out["step_ms_back_to_back"] = a.elapsed_time(e) / reps
    # pinned host features through the stream, eager and replayed
    xh = x.cpu().pin_memory()
    for replay in (False, True):
        gs = m.GraphStream(net, dev, depth=2, graph_replay=replay)
        preds = [torch.empty(ei.shape[1], dtype=torch.uint8).pin_memory() for _ in range(2)]
        for i in range(6):
            gs.submit(xh, cam_host, preds[i & 1])
        gs.drain()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(50):
            gs.submit(xh, cam_host, preds[i & 1])
        gs.drain()
        e.record()
        e.synchronize()
        out["stream_ms_replay" if replay else "stream_ms_eager"] = a.elapsed_time(e) / 50
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/step_gaps_%s.json" % out["balance"], "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
