"""N>1 path on CPU: row-block partitioning + the collective schedule, with gloo (world_size 2) and in-process shards."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gcn_mtmc_b200 as m
from oracle import mpn_oracle as mo
from tests.fake_phases import FakePhases, FakePostOps


def _case(L, n_cls, N=36, C=3, D=48):
    params = mo.shipped_model_params(L, n_cls, D, (40,))
    x, ei, cam, _ = mo.synth_graph(N, C, 5, D=D, planted=True)
    sd = mo.init_weights(params, "resnet101", 11)
    ea = mo.edge_features(x, ei)
    ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    return params, sd, x, ei, ea, ref, href


def _blocks(ei, N, world):
    rowptr = torch.searchsorted(ei[0].contiguous(), torch.arange(N + 1))
    return m.partition_rows(rowptr, world)


class NoComm:
    world, rank = 1, 0

    def all_reduce_sum(self, t):
        pass

    def all_gather_rows(self, full, blocks):
        pass

    def all_gather_ragged(self, t):
        return [t]


@pytest.mark.parametrize("L,n_cls,world", [(1, 1, 1), (1, 1, 3), (3, 2, 2), (4, 4, 4), (0, 1, 2)])
def test_inprocess_shards_match_oracle(L, n_cls, world):
    params, sd, x, ei, ea, ref, href = _case(L, n_cls)
    N = x.shape[0]
    blocks = _blocks(ei, N, world)
    assert blocks[0][0] == 0 and blocks[-1][1] == N and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
    n_out = 1 if L == 0 else n_cls
    shards, spans = [], []
    for (n0, n1) in blocks:
        lo, hi = m.shard_edges(ei, n0, n1)
        spans.append((lo, hi))
        shards.append(FakePhases(sd, params, x, ei[:, lo:hi], ea[lo:hi], n0, n1, ei.shape[1], n_out))
    assert spans[0][0] == 0 and spans[-1][1] == ei.shape[1]
    k = m.sharded_forward(shards, NoComm(), L, n_cls, blocks)
    assert k == n_out
    for i in range(n_out):
        got = torch.cat([s.logits[i] for s in shards])
        assert (got - ref[i]).abs().max().item() <= 1e-9
    h = torch.cat([s.h_full()[b[0]:b[1]] for s, b in zip(shards, blocks)])
    assert (h - href).abs().max().item() <= 1e-9


def _worker(rank, world, port, L, n_cls, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params, sd, x, ei, ea, ref, href = _case(L, n_cls)
        blocks = _blocks(ei, x.shape[0], world)
        n0, n1 = blocks[rank]
        lo, hi = m.shard_edges(ei, n0, n1)
        tot = torch.tensor([hi - lo], dtype=torch.float64)
        dist.all_reduce(tot)
        assert int(tot.item()) == ei.shape[1]
        ph = FakePhases(sd, params, x, ei[:, lo:hi], ea[lo:hi], n0, n1, int(tot.item()), n_cls)
        m.sharded_forward(ph, m.sharded.TorchComm(), L, n_cls, blocks)
        err = max((ph.logits[i] - ref[i][lo:hi]).abs().max().item() for i in range(n_cls))
        herr = (ph.h_full()[n0:n1] - href[n0:n1]).abs().max().item()
        q.put((rank, err, herr))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("L,n_cls", [(1, 1), (3, 2)])
def test_gloo_world2_matches_oracle(L, n_cls):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, L, n_cls, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] <= 1e-9 and r[2] <= 1e-9 for r in res), res


# ---------------------------------------------------------------------------------------------------------------
# sharded post-processing: shard compaction -> one exchange of the active lists -> rounds on the merged list
# ---------------------------------------------------------------------------------------------------------------
def _post_case(seed=3, n_nodes=400, cams=5):
    import numpy as np
    from oracle import postproc_oracle as po
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, seed, n_extra_per_node=5.0, flip_on=0.06,
                                                          flip_off=0.04, single_dir=0.06)
    order = np.lexsort((dst, src))
    return src[order], dst[order], prob[order], pred[order], n_nodes, cams


def _post_shards(src, dst, prob, pred, n_nodes, world):
    from types import SimpleNamespace
    ei = torch.from_numpy(__import__("numpy").stack([src, dst]))
    rowptr = torch.searchsorted(ei[0].contiguous(), torch.arange(n_nodes + 1))
    out = []
    for (n0, n1) in m.partition_rows(rowptr, world):
        lo, hi = m.shard_edges(ei, n0, n1)
        out.append((lo, hi, (SimpleNamespace(src=ei[0, lo:hi], dst=ei[1, lo:hi], n_edges=hi - lo, n_cols=n_nodes),
                             torch.from_numpy(pred[lo:hi].astype("uint8")), torch.from_numpy(prob[lo:hi]))))
    return out


FLAG_SETS = [(True, True, True), (True, False, False), (False, True, False), (False, False, True), (False, False, False)]


@pytest.mark.parametrize("world", [1, 3])
@pytest.mark.parametrize("flags", FLAG_SETS)
def test_sharded_post_processing_inprocess_matches_whole_graph(world, flags):
    import numpy as np
    from oracle import postproc_oracle as po
    src, dst, prob, pred, n_nodes, cams = _post_case()
    ref_lab, ref_act = po.post_processing_rounds(src, dst, pred, prob, cams, n_nodes, cutting=flags[0], pruning=flags[1],
                                                 splitting=flags[2], numbering="reference")
    sh = _post_shards(src, dst, prob, pred, n_nodes, world)
    CONFIG = {"CUTTING": flags[0], "PRUNING": str(flags[1]), "SPLITTING": flags[2]}
    ID, preds = m.sharded_post_processing(cams, [t for (_lo, _hi, t) in sh], CONFIG, n_nodes, comm=NoComm(), ops=FakePostOps())
    assert np.array_equal(torch.cat(preds).numpy().astype(np.int64), ref_act)
    assert np.array_equal(ID.numpy(), ref_lab)
    assert all(p.data_ptr() == t[1].data_ptr() for p, (_lo, _hi, t) in zip(preds, sh))      # updated in place


def test_sharded_post_processing_no_active_edges():
    src, dst, prob, pred, n_nodes, cams = _post_case()
    sh = _post_shards(src, dst, prob, 0 * pred, n_nodes, 2)
    ID, preds = m.sharded_post_processing(cams, [t for (_lo, _hi, t) in sh], {"CUTTING": True, "PRUNING": True, "SPLITTING": True},
                                          n_nodes, comm=NoComm(), ops=FakePostOps())
    assert ID.tolist() == list(range(n_nodes)) and not any(bool(p.any()) for p in preds)


def test_sharded_post_processing_default_comm_without_process_group():
    """comm=None outside torch.distributed: a world of one, the whole graph as a single shard."""
    import numpy as np
    from oracle import postproc_oracle as po
    src, dst, prob, pred, n_nodes, cams = _post_case(seed=11)
    ref_lab, ref_act = po.post_processing_rounds(src, dst, pred, prob, cams, n_nodes, numbering="reference")
    (_lo, _hi, triple), = _post_shards(src, dst, prob, pred, n_nodes, 1)
    ID, p = m.sharded_post_processing(cams, triple, {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, n_nodes, ops=FakePostOps())
    assert np.array_equal(p.numpy().astype(np.int64), ref_act) and np.array_equal(ID.numpy(), ref_lab)


def _post_worker(rank, world, port, q, empty_rank=-1):
    import numpy as np
    from oracle import postproc_oracle as po
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        src, dst, prob, pred, n_nodes, cams = _post_case(seed=7)
        if empty_rank >= 0:                                # one rank contributes an EMPTY active list to the ragged all-gather
            lo, hi, _ = _post_shards(src, dst, prob, pred, n_nodes, world)[empty_rank]
            pred = pred.copy()
            pred[lo:hi] = 0
        ref_lab, ref_act = po.post_processing_rounds(src, dst, pred, prob, cams, n_nodes, numbering="reference")
        lo, hi, triple = _post_shards(src, dst, prob, pred, n_nodes, world)[rank]
        ID, p = m.sharded_post_processing(cams, triple, {"CUTTING": "True", "PRUNING": True, "SPLITTING": True}, n_nodes,
                                          comm=m.sharded.TorchComm(), ops=FakePostOps())
        q.put((rank, bool(np.array_equal(p.numpy().astype(np.int64), ref_act[lo:hi])), bool(np.array_equal(ID.numpy(), ref_lab)),
               int(ref_act.sum()) < int(pred.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("empty_rank", [-1, 1])
def test_sharded_post_processing_gloo_world2(empty_rank):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (7 if empty_rank >= 0 else 0)
    procs = [ctx.Process(target=_post_worker, args=(r, 2, port, q, empty_rank)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted(q.get(timeout=5) for _ in range(2))
    assert res == [(0, True, True, True), (1, True, True, True)], res
