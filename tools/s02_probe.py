"""S02-shaped graph through GraphStream(depth=1): eager vs replay, and whether the capture happened."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
import gcn_mtmc_b200 as m
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
net = bench.make_model(dev)
x, ei = bench.device_graph(300, 4, 0, dev)
hx = x.cpu().pin_memory()
hp = torch.empty(ei.shape[1], dtype=torch.uint8).pin_memory()
cam_host = (torch.arange(300) * 4 // 300).numpy()
for replay in (False, True, False, True):
    gs = m.GraphStream(net, dev, depth=1, graph_replay=replay)
    for _ in range(6):
        gs.submit(hx, cam_host, hp); gs.drain()
    ts, hs = [], []
    for _ in range(200):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record(); gs.submit(hx, cam_host, hp); t1 = time.perf_counter(); gs.drain(host_sync=False); e.record(); e.synchronize()
        hs.append(1e3 * (t1 - t0)); ts.append(a.elapsed_time(e))
    ts.sort(); hs.sort()
    print("replay", replay, "captured", gs.slots[0].cap is not None, "p50 %.4f ms" % ts[100], "host submit p50 %.4f ms" % hs[100])
