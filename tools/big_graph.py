"""BASELINE configs[4]: L=4 message-passing steps on a 32k-tracklet / 8-camera graph (E = 939,524,096 directed edges),
row-block sharded over the GPUs of one box (strong scaling: the graph is fixed, every rank owns N/world rows).

    python tools/big_graph.py [--nodes 32768] [--enc-steps 4] [--steps 5]                       # one GPU
    python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tools/big_graph.py ...

The graph tables are built on the device from the camera ids (no int64 edge_index: 15 GB at this size).  A step = K0 (tables)
+ K1 (edge features of the rank's rows) + forward with fused decisions.  Prints one JSON line (rank 0): directed edges/s,
edge-steps/s, device ms per step (CUDA events, max over ranks).  Diagnostic numbers for DESIGN.md / profiles, not the bench line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=32768)
    ap.add_argument("--cams", type=int, default=8)
    ap.add_argument("--enc-steps", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import gcn_mtmc_b200 as m
    m._lib.require_device(local_rank)
    net = bench.make_model(dev, L=args.enc_steps, n_cls=1)
    N, Cn = args.nodes, args.cams
    assert N % (Cn * world) == 0
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.nn.functional.normalize(torch.randn(N, bench.FEAT_DIM, generator=g, device=dev), p=2, dim=0)
    cam_host = (torch.arange(N) * Cn // N).numpy()
    per = N // world
    blocks = [(r * per, (r + 1) * per) for r in range(world)]
    E_total = N * (N - N // Cn)
    sharded = m.ShardedMPN(net) if world > 1 else None
    batch = bench.Batch()
    batch.num_nodes = N

    def step():
        if world == 1:
            gr = m.TrackletGraph.from_cameras(cam_host, dev)
            batch.x, batch.mpn_graph = x, gr
            batch.edge_attr = m.edge_features(x, None, graph=gr)
            net(batch)
            return net.last_pred
        gr = m.TrackletGraph.from_cameras(cam_host, dev, row_block=blocks[rank])
        ea = m.edge_features(x, None, graph=gr)
        return sharded.forward(x, None, ea, blocks, fuse_decisions=True, graph=gr, total_edges=E_total)[2]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    tot = 0.0
    for _ in range(args.steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        pred = step()
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        tot += float(ms.item())
    ms = tot / args.steps
    active = torch.tensor([float(pred.sum().item())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(active)
    if rank == 0:
        print(json.dumps({"config": "N=%d C=%d L=%d E=%d directed edges, row-block sharded over %d GPU(s), tables from camera ids" %
                                    (N, Cn, args.enc_steps, E_total, world),
                          "n_gpus": world, "ms_per_step": ms, "edges_per_s": E_total / (ms * 1e-3),
                          "edge_steps_per_s": E_total * max(args.enc_steps, 1) / (ms * 1e-3), "active_edges": int(active.item()),
                          "peak_mem_gb_rank0": torch.cuda.max_memory_allocated(dev) / 1e9, "fused_peer_path": bool(sharded.fused) if sharded else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
