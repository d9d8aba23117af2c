"""Host-side cost of enqueueing one configs[1] step (no synchronisation inside the timed region): how long the Python + ctypes +
launch path takes for the graph tables and for the forward call, against the device time of the same step.

    python tools/host_enqueue.py          # gpurun_out/host_enqueue.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gcn_mtmc_b200 as m


def main():
    N, cams = bench.NODES_1GPU, bench.CAMS
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m._lib.require_device(0)
    x, ei = bench.device_graph(N, cams, 0, dev)
    cam_host = (torch.arange(N) * cams // N).numpy()
    net = bench.make_model(dev)
    b = bench.Batch()
    out = {}
    for name, mk in (("edge_index", lambda: m.TrackletGraph(ei, N, validate="deferred")),
                     ("cameras", lambda: m.TrackletGraph.from_cameras(cam_host, dev))):
        tg, tf, td = [], [], []
        for i in range(30):
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            g = mk()
            t1 = time.perf_counter()
            b.x, b.edge_index, b.num_nodes, b.edge_attr, b.mpn_graph = x, ei, N, None, g
            net(b)
            t2 = time.perf_counter()
            e.record()
            e.synchronize()
            if i >= 5:
                tg.append(1e3 * (t1 - t0)); tf.append(1e3 * (t2 - t1)); td.append(a.elapsed_time(e))
        med = lambda v: sorted(v)[len(v) // 2]
        out[name] = {"host_ms_graph_tables": med(tg), "host_ms_forward_call": med(tf), "device_ms_step": med(td)}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/host_enqueue.json", "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
