"""Multi-GPU parity check, launched by torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py

Each rank runs its row-block shard through ShardedMPN, once with the collectives fused into the kernels over NVLink
peer memory (mpn_forward_sharded) and once with the NCCL schedule; both are compared with the fp64 oracle of the whole
graph.  Repeated calls exercise the sequence-number / slot reuse of the peer protocol.  The decisions then go through
the sharded post-processing (active lists exchanged over NCCL) and are compared with the CPU restatement on the whole graph.
"""
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gcn_mtmc_b200 as m                      # noqa: E402
from oracle import mpn_oracle as mo            # noqa: E402
from oracle import postproc_oracle as po       # noqa: E402


def check_case(rank, world, dev, L, n_cls, N, C, modes, chunk=None, reattach=False):
    params = mo.shipped_model_params(L, n_cls, 128, (96, 64))
    params["reattach_initial_nodes"] = params["reattach_initial_edges"] = reattach
    x, ei, cam, _ = mo.synth_graph(N, C, 3, D=128, planted=True)
    sd = mo.init_weights(params, "resnet101", 2)
    ea = mo.edge_features(x, ei)
    ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    rowptr = torch.searchsorted(ei[0].contiguous(), torch.arange(N + 1))
    blocks = m.partition_rows(rowptr, world)
    n0, n1 = blocks[rank]
    lo, hi = m.shard_edges(ei, n0, n1)
    xd = x.to(dev)
    ei_l = ei[:, lo:hi].to(dev)
    g = m.TrackletGraph(ei_l, N, row_offset=n0, n_rows=n1 - n0, chunk=chunk)      # chunk >= 128: tensor-core apply kernel
    ea_l = m.edge_features(xd, None, graph=g, use_tensor_cores=False)
    assert torch.allclose(ea_l.cpu(), ea[lo:hi], rtol=3e-6, atol=3e-6)
    worst = 0.0
    results = {}
    for fused in (True, False):
        sh = m.ShardedMPN(net, fused=fused)
        for _rep in range(3):
            out, h_l, pred, prob1 = sh.forward(xd, ei_l, ea[lo:hi].to(dev), blocks, fuse_decisions=True, graph=g)
        if fused:                                      # edge features inside the call (fused kernel + its moment sums): same logits
            out_f, h_f, pred_f, _ = sh.forward(xd, ei_l, None, blocks, fuse_decisions=True, graph=g)
            assert torch.allclose(sh.last_edge_attr.cpu(), ea[lo:hi], rtol=1e-5, atol=1e-5)
            assert sh.shared_gram_used(), "the shared symmetric Gram did not run (dense cross-camera rows on every rank)"
            scale = ref[-1].abs().max().item()
            assert (out_f["classified_edges"][-1] - out["classified_edges"][-1]).abs().max().item() <= 1e-5 * scale
            for _rep in range(2):                      # sequence numbers of the node-table exchange, buffer reuse
                out_g, _, pred_g, _ = sh.forward(xd, ei_l, None, blocks, fuse_decisions=True, graph=g)
                assert torch.equal(out_g["classified_edges"][-1], out_f["classified_edges"][-1]) and torch.equal(pred_g, pred_f)
            sh_own = m.ShardedMPN(net, fused=True, shared_gram=False)      # every rank all pairs of its own rows
            out_o, _, pred_o, _ = sh_own.forward(xd, ei_l, None, blocks, fuse_decisions=True, graph=g)
            assert not sh_own.shared_gram_used()
            assert torch.allclose(sh_own.last_edge_attr.cpu(), ea[lo:hi], rtol=1e-5, atol=1e-5)
            assert (out_o["classified_edges"][-1] - out_f["classified_edges"][-1]).abs().max().item() <= 1e-5 * scale
        torch.cuda.synchronize()
        if fused and sh.path != "fused_peer_memory" and rank == 0:
            print("fused path unavailable")
        mode = "fused" if (fused and sh.peers is not None) else "nccl"
        modes.add(mode)
        results[mode] = out["classified_edges"][-1].clone()
        for i in range(n_cls):
            err = (out["classified_edges"][i].cpu().double() - ref[i][lo:hi]).abs().max().item()
            tol = 1e-4 * ref[i].abs().max().item()
            assert err <= tol, (mode, L, i, err, tol)
            worst = max(worst, err / tol)
        assert (h_l.cpu().double() - href[n0:n1]).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item()), mode
        margin = (ref[-1][lo:hi, 1] - ref[-1][lo:hi, 0]).abs()
        bad = (pred.cpu().long() != ref[-1][lo:hi].argmax(1)) & (margin > 1e-4)
        assert not bool(bad.any()), mode
    if len(results) == 2:
        assert (results["fused"] - results["nccl"]).abs().max().item() <= 2e-6
    # the decisions of the last forward go straight into the sharded post-processing (shard compaction, one exchange of the
    # active lists over NCCL, rounds on the merged list): bit-exact against the CPU restatement run on the gathered graph
    parts = [None] * world
    dist.all_gather_object(parts, (pred.cpu().numpy(), prob1.cpu().numpy()))
    pred_all = np.concatenate([p[0] for p in parts]).astype(np.int64)
    prob_all = np.concatenate([p[1] for p in parts])
    lab_ref, act_ref = po.post_processing_rounds(ei[0].numpy(), ei[1].numpy(), pred_all, prob_all, C, N, numbering="reference")
    ID, new_pred = sh.post_processing(C, g, pred, prob1, {"CUTTING": True, "PRUNING": "True", "SPLITTING": True})
    assert new_pred.data_ptr() == pred.data_ptr()
    assert np.array_equal(new_pred.cpu().numpy().astype(np.int64), act_ref[lo:hi]), "sharded post-processing: decisions"
    assert np.array_equal(ID.numpy(), lab_ref), "sharded post-processing: labels"
    return worst


def check_shared_fallback(rank, world, dev, N=360, C=4):
    """One row of the LAST rank's block lacks an edge (not a dense cross-camera row): that rank's rows go through the Gram + gather
    path, and every other rank must notice — through the node-table exchange, on the device — and compute all pairs of its own rows
    instead of waiting for mirrored entries that never come.  Features and logits against the oracle of the modified graph."""
    params = mo.shipped_model_params(1, 1, 128, (96, 64))
    x, ei, cam, _ = mo.synth_graph(N, C, 7, D=128, planted=True)
    victim = N - 2                                               # a row of the last block: drop one of its middle edges
    rows = np.flatnonzero(ei[0].numpy() == victim)
    keep = np.ones(ei.shape[1], dtype=bool)
    keep[rows[len(rows) // 2]] = False
    ei = ei[:, torch.from_numpy(keep)]
    sd = mo.init_weights(params, "resnet101", 5)
    ea = mo.edge_features(x, ei)
    ref, _ = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    rowptr = torch.searchsorted(ei[0].contiguous(), torch.arange(N + 1))
    blocks = m.partition_rows(rowptr, world)
    n0, n1 = blocks[rank]
    assert blocks[-1][0] <= victim
    lo, hi = m.shard_edges(ei, n0, n1)
    ei_l = ei[:, lo:hi].to(dev)
    g = m.TrackletGraph(ei_l, N, row_offset=n0, n_rows=n1 - n0)
    sh = m.ShardedMPN(net)
    for _rep in range(2):
        out, _, pred, _ = sh.forward(x.to(dev), ei_l, None, blocks, fuse_decisions=True, graph=g)
        assert not sh.shared_gram_used(), "a rank used the shared Gram although one rank's rows are not dense cross-camera rows"
        assert torch.allclose(sh.last_edge_attr.cpu(), ea[lo:hi], rtol=1e-5, atol=1e-5)
        err = (out["classified_edges"][-1].cpu().double() - ref[-1][lo:hi]).abs().max().item()
        assert err <= 1e-4 * ref[-1].abs().max().item(), err
    # and the shared mode comes back for the next dense graph on the same ShardedMPN / exchange buffers
    x2, ei2, _, _ = mo.synth_graph(N, C, 8, D=128, planted=True)
    rowptr2 = torch.searchsorted(ei2[0].contiguous(), torch.arange(N + 1))
    blocks2 = m.partition_rows(rowptr2, world)
    lo2, hi2 = m.shard_edges(ei2, *blocks2[rank])
    g2 = m.TrackletGraph(ei2[:, lo2:hi2].to(dev), N, row_offset=blocks2[rank][0], n_rows=blocks2[rank][1] - blocks2[rank][0])
    sh.forward(x2.to(dev), ei2[:, lo2:hi2].to(dev), None, blocks2, fuse_decisions=True, graph=g2)
    assert sh.shared_gram_used()
    assert torch.allclose(sh.last_edge_attr.cpu(), mo.edge_features(x2, ei2)[lo2:hi2], rtol=1e-5, atol=1e-5)


def check_post(rank, world, dev, n_nodes=20000, cams=8):
    """Sparse predicted graph (planted clusters + noise), sharded by row block: decisions and reference label integers."""
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, 9, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03,
                                                          single_dir=0.05)
    order = np.lexsort((dst, src))
    src, dst, prob, pred = src[order], dst[order], prob[order], pred[order]
    lab_ref, act_ref = po.post_processing_rounds(src, dst, pred, prob, cams, n_nodes, numbering="reference")
    ei = torch.from_numpy(np.stack([src, dst]))
    rowptr = torch.searchsorted(ei[0].contiguous(), torch.arange(n_nodes + 1))
    n0, n1 = m.partition_rows(rowptr, world)[rank]
    lo, hi = m.shard_edges(ei, n0, n1)
    g = m.TrackletGraph(ei[:, lo:hi].to(dev), n_nodes, row_offset=n0, n_rows=n1 - n0)
    p_l = torch.from_numpy(pred[lo:hi].astype(np.uint8)).to(dev)
    q_l = torch.from_numpy(prob[lo:hi]).to(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    ID, p_new = m.sharded_post_processing(cams, (g, p_l, q_l), {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, n_nodes)
    ev1.record()
    torch.cuda.synchronize()
    assert np.array_equal(p_new.cpu().numpy().astype(np.int64), act_ref[lo:hi]) and np.array_equal(ID.numpy(), lab_ref)
    assert int(act_ref.sum()) < int(pred.sum())
    return ev0.elapsed_time(ev1)


def check_stream(rank, world, dev):
    """ShardedGraphStream (each rank copies its own feature rows, NVLink all-gather, shard forward, decisions back; two graphs in
    flight) against the direct sharded call, bit for bit, over several graphs of two shapes."""
    params = mo.shipped_model_params(2, 1, 64, (48, 40))
    sd = mo.init_weights(params, "resnet101", 3)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    sh = m.ShardedMPN(net)
    for N, C in ((64 * world * 3, 3), (128 * world * 4, 4)):
        per = N // world
        blocks = [(r * per, (r + 1) * per) for r in range(world)]
        cam = (np.arange(N) * C // N)
        xs, refs = [], []
        for seed in range(5):
            x, _, _, _ = mo.synth_graph(N, C, seed, D=64, planted=True)
            xd = x.to(dev)
            g = m.TrackletGraph.from_cameras(cam, dev, row_block=blocks[rank])
            pred = sh.forward(xd, None, None, blocks, fuse_decisions=True, graph=g)[2]      # the call the stream makes
            xs.append(x[blocks[rank][0]:blocks[rank][1]].contiguous().pin_memory())
            refs.append(pred.cpu())
        gs = m.ShardedGraphStream(sh, blocks, dev, depth=2)
        outs = [torch.zeros(refs[0].numel(), dtype=torch.uint8).pin_memory() for _ in xs]
        for xr, o in zip(xs, outs):
            gs.submit(xr, cam, o)
        gs.drain()
        for o, r in zip(outs, refs):
            assert torch.equal(o, r), "ShardedGraphStream decisions differ from the direct sharded call"


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    worst, modes = 0.0, set()
    for (L, n_cls, N, C, chunk, reattach) in [(1, 1, 240, 4, None, False), (4, 2, 200, 5, None, False), (2, 1, 600, 3, 128, False),
                                              (1, 1, 600, 3, 256, False), (3, 1, 600, 3, 128, True)]:
        worst = max(worst, check_case(rank, world, dev, L, n_cls, N, C, modes, chunk, reattach))
    check_shared_fallback(rank, world, dev)
    post_ms = check_post(rank, world, dev)
    check_stream(rank, world, dev)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("dist_gpu_check ok: world=%d worst err/tol=%.3f modes=%s sharded post-processing (20000 nodes) %.2f ms"
              % (world, t.item(), sorted(modes), post_ms))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
