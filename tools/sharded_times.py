"""Per-step device times of the sharded forward, fused (peer memory) vs NCCL schedule.  Launch with torchrun."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import gcn_mtmc_b200 as m

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
net = bench.make_model(dev)
n_nodes = int(round(bench.NODES_1GPU * world ** 0.5 / (bench.CAMS * world))) * bench.CAMS * world
per = n_nodes // world
blocks = [(r * per, (r + 1) * per) for r in range(world)]
x, ei = bench.device_graph(n_nodes, bench.CAMS, 0, dev, row_block=blocks[rank])
n0, n1 = blocks[rank]
g = m.TrackletGraph(ei, n_nodes, row_offset=n0, n_rows=n1 - n0)
ea = m.edge_features(x, ei, graph=g)
for fused in (True, False):
    sh = m.ShardedMPN(net, fused=fused)
    ts, hs = [], []
    for i in range(12):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record(); sh.forward(x, ei, ea, blocks, fuse_decisions=True, graph=g); b.record(); b.synchronize()
        hs.append(1e3 * (time.perf_counter() - t0)); ts.append(a.elapsed_time(b))
    if rank == 0:
        print("fused=%s peers=%s path=%s" % (fused, sh.peers is not None, sh.path))
        print("  device ms:", " ".join("%.2f" % t for t in ts))
        print("  host   ms:", " ".join("%.2f" % t for t in hs))
tt = torch.tensor([ei.shape[1]], dtype=torch.float64, device=dev); dist.all_reduce(tt); TOT = int(tt.item())
for fused in (True, False):
    sh = m.ShardedMPN(net, fused=fused)
    ts, parts = [], []
    for i in range(10):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); a.record()
        g2 = m.TrackletGraph(ei, n_nodes, row_offset=n0, n_rows=n1 - n0); torch.cuda.synchronize(); t1 = time.perf_counter()
        ea2 = m.edge_features(x, ei, graph=g2); torch.cuda.synchronize(); t2 = time.perf_counter()
        sh.forward(x, ei, ea2, blocks, fuse_decisions=True, graph=g2); b.record(); b.synchronize(); t3 = time.perf_counter()
        ts.append(a.elapsed_time(b)); parts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
    if rank == 0:
        print("bench-like fused=%s" % fused)
        print("  device ms:", " ".join("%.2f" % t for t in ts))
        print("  graph/ef/fwd host ms:", " ".join("%.2f/%.2f/%.2f" % p for p in parts[-4:]))
dist.destroy_process_group()
