"""CUDA post-processing (through the C ABI) against the plain-C oracle at a size the Python restatement needs minutes for.
B200 only.  Named to run last: it was added after the round's GPU time was spent, and the same input is checked on the CPU
in tests/test_c_oracle.py (C oracle == Python rounds oracle, same generator arguments)."""
import numpy as np
import pytest
import torch

from oracle import postproc_c as pc
from oracle import postproc_oracle as po
from tests._util import Data, same_partition

pytestmark = pytest.mark.gpu

# the same arguments as tests/test_c_oracle.py::test_c_oracle_pins_the_gpu_case
GPU_CASE = dict(n_nodes=200000, cams=8, seed=3, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03, single_dir=0.05)


@pytest.fixture(scope="module")
def m():
    import gcn_mtmc_b200 as mod
    mod._lib.require_device(0)
    return mod


def test_post_processing_200k_nodes_vs_c_oracle(m):
    """200 k nodes, 2.2 M edges, ~0.9 M active when SPLITTING starts, 1568 tied probability values: the default path against the
    reference's own SPLITTING order (oracle/postproc_oracle.c::po_split_sequential, stored by tests/golden/make_split200k.py —
    8 minutes of CPU) with ZERO differing decisions, and the reference's label integers; CUT / PRUNE against the C oracle live."""
    import hashlib
    import os
    from tests._util import GOLDEN
    c = GPU_CASE
    dev = torch.device("cuda", 0)
    src, dst, prob, pred, _ = po.planted_prediction_graph(c["n_nodes"], c["cams"], c["seed"], n_extra_per_node=c["n_extra_per_node"],
                                                          flip_on=c["flip_on"], flip_off=c["flip_off"], single_dir=c["single_dir"])
    gold = np.load(os.path.join(GOLDEN, "split200k_sequential.npz"))
    assert list(gold["spec"]) == [c["n_nodes"], c["cams"], c["seed"]] and int(gold["n_edges"][0]) == src.size
    start = pc.cut(src, dst, pred, c["n_nodes"])
    start, _ = pc.prune(src, dst, start, prob, c["cams"], c["n_nodes"])
    start = pc.cut(src, dst, start, c["n_nodes"])
    start_idx = np.flatnonzero(start)
    assert start_idx.size == int(gold["n_start_active"][0])
    act_ref = np.zeros(src.size, dtype=np.int64)
    act_ref[start_idx] = np.unpackbits(gold["final_bits"])[:start_idx.size]
    data = Data(x=torch.zeros(c["n_nodes"], 1, device=dev), edge_index=torch.from_numpy(np.stack([src, dst])).to(dev))
    cfg = {"CUTTING": True, "PRUNING": True, "SPLITTING": True}
    ID, P = m.post_processing(c["cams"], None, None, torch.from_numpy(pred).to(dev), None, cfg, data, torch.from_numpy(prob).to(dev))
    stats = m.split_stats()
    assert stats["mode"] == "reference_order_host" and stats["tied_edges"] > 0 and stats["rounds"] > 0
    assert int((P.cpu().numpy() != act_ref).sum()) == 0                            # decisions: bit-exact, ties included
    assert hashlib.sha256(np.ascontiguousarray(ID.numpy(), dtype=np.int64).tobytes()).digest() == gold["labels_sha256"].tobytes()
    assert np.bincount(ID.numpy()).max() <= c["cams"]                              # size-independent property
    IDc, Pc = m.post_processing(c["cams"], None, None, torch.from_numpy(pred).to(dev), None, cfg, data, torch.from_numpy(prob).to(dev),
                                numbering="canonical")
    assert np.array_equal(Pc.cpu().numpy(), act_ref) and same_partition(IDc.numpy(), ID.numpy())
    # the rounds formulation (what round 1 shipped) differs on this graph: the fixture does pin the order
    assert int((pc.split(src, dst, start, prob, c["cams"], c["n_nodes"]) != act_ref).sum()) > 0


def test_splitting_under_ties_matches_the_reference(m):
    """The DEFAULT post_processing against the outputs of the UNMODIFIED reference on graphs with exact probability ties
    (tests/golden/ties_cases.npz; on half of them all-clusters-per-round gives other decisions), three flag sets each; and graphs
    without ties take the device rounds."""
    from tests.test_c_oracle import _ties_cases
    dev = torch.device("cuda", 0)
    host_mode = 0
    for i, s, d, prob, pred, C, n, ref in _ties_cases():
        data = Data(x=torch.zeros(n, 1, device=dev), edge_index=torch.from_numpy(np.stack([s, d])).to(dev))
        for tag, cfg in (("full", (True, True, True)), ("split_only", (False, False, True)), ("prune_split", (False, True, True))):
            CONFIG = {"CUTTING": cfg[0], "PRUNING": cfg[1], "SPLITTING": cfg[2]}
            ID, P = m.post_processing(C, None, None, torch.from_numpy(pred).to(dev), None, dict(CONFIG), data,
                                      torch.from_numpy(prob).to(dev))
            host_mode += m.split_stats()["mode"] == "reference_order_host"
            assert np.array_equal(P.cpu().numpy(), ref["pred_" + tag]), (i, tag)          # the reference's decisions
            assert np.array_equal(ID.numpy(), ref["labels_" + tag]), (i, tag)             # and its label integers
            if tag == "split_only":                                                       # splitting() mutates its argument (utils.py:98)
                p2 = torch.from_numpy(pred).to(dev)
                out = m.splitting(None, p2, torch.from_numpy(prob).to(dev), None, data, None, C)
                assert out is p2 and np.array_equal(p2.cpu().numpy(), ref["pred_" + tag]), (i, tag)
    assert host_mode >= 10
    n_nodes, cams = 20000, 8
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, 2, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03,
                                                          single_dir=0.05)
    start = pc.cut(src, dst, pred, n_nodes)
    start, _ = pc.prune(src, dst, start, prob, cams, n_nodes)
    start = pc.cut(src, dst, start, n_nodes)
    act_ref = pc.split_sequential(src, dst, start, prob, cams, n_nodes)                   # the reference's order (5 s of CPU)
    lab_ref, _ = pc.scc_labels(src, dst, act_ref, n_nodes)
    data = Data(x=torch.zeros(n_nodes, 1, device=dev), edge_index=torch.from_numpy(np.stack([src, dst])).to(dev))
    for numbering in ("reference", "canonical"):
        ID, P = m.post_processing(cams, None, None, torch.from_numpy(pred).to(dev), None,
                                  {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, data, torch.from_numpy(prob).to(dev),
                                  numbering=numbering)
        assert np.array_equal(P.cpu().numpy(), act_ref)
        assert np.array_equal(ID.numpy(), lab_ref) if numbering == "reference" else same_partition(ID.numpy(), lab_ref)
    # distinct probabilities -> no tie -> device rounds, same result as the reference's order
    rng = np.random.default_rng(0)
    prob_u = rng.permutation(np.linspace(0.55, 0.99, prob.size, dtype=np.float64)).astype(np.float32)
    prob_u = np.where(pred > 0, prob_u, np.float32(1.0) - prob_u).astype(np.float32)
    assert np.unique(prob_u[pred > 0]).size == int((pred > 0).sum())
    ID, P = m.post_processing(cams, None, None, torch.from_numpy(pred).to(dev), None,
                              {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, data, torch.from_numpy(prob_u).to(dev))
    st = m.split_stats()
    assert st["mode"] == "device_rounds" and st["tied_edges"] == 0 and st["rounds"] > 0
    start = pc.cut(src, dst, pred, n_nodes)
    start, _ = pc.prune(src, dst, start, prob_u, cams, n_nodes)
    start = pc.cut(src, dst, start, n_nodes)
    assert np.array_equal(P.cpu().numpy(), pc.split_sequential(src, dst, start, prob_u, cams, n_nodes))


def test_post_processing_matches_the_reference_at_s02_size(m):
    """The CUDA post-processing against the unmodified reference's outputs on one S02-shaped predicted graph (300 tracklets, 4 cameras,
    67,500 directed edges: the reference's own problem size; tests/golden/s02post_n300_c4.npz), all five flag sets."""
    from tests.test_c_oracle import S02_FLAG_SETS, load_s02_case
    g, N, C, src, dst, prob, pred = load_s02_case()
    dev = torch.device("cuda", 0)
    data = Data(x=torch.zeros(N, 1, device=dev), edge_index=torch.from_numpy(np.stack([src, dst])).to(dev))
    for tag, cfg in S02_FLAG_SETS:
        CONFIG = {"CUTTING": cfg[0], "PRUNING": cfg[1], "SPLITTING": cfg[2]}
        ID, P = m.post_processing(C, None, None, torch.from_numpy(pred).to(dev), None, CONFIG, data, torch.from_numpy(prob).to(dev))
        assert np.array_equal(P.cpu().numpy(), g["pred_" + tag]), tag
        assert np.array_equal(ID.numpy(), g["labels_" + tag]), tag


def test_forward_matches_the_reference_at_s02_size(m):
    """The CUDA edge features, and forward + decisions on the reference's edge features, against the unmodified reference on one S02-shaped graph (300 tracklets, 4
    cameras, 2048-d features, 67,500 directed edges, shipped L=1 model; tests/golden/s02mpn_shipped_L1.npz): logits within
    1e-4 * max|logit| (north_star), h within 1e-4, decisions identical outside the 1e-4 margin band, edge features within 1e-5."""
    import copy
    import os
    from tests._util import GOLDEN, load_mpn_case
    dev = torch.device("cuda", 0)
    g, params, sd, x, ei, _ = load_mpn_case(os.path.join(GOLDEN, "s02mpn_shipped_L1.npz"))
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev).eval()
    net.fuse_decisions = True
    feats = m.edge_features(x.to(dev), ei.to(dev))                               # inference.py:453-456 on the CUDA path
    assert np.allclose(feats.cpu().numpy(), g["edge_attr"], rtol=1e-5, atol=1e-5)
    data = Data(x=x.to(dev), edge_index=ei.to(dev), edge_attr=torch.from_numpy(g["edge_attr"]).to(dev))   # the reference's own input
    out, h = net(data)
    torch.cuda.synchronize()
    ref = g["logits0"]
    logits = out["classified_edges"][-1].cpu().numpy()
    assert len(out["classified_edges"]) == 1 and np.abs(logits - ref).max() <= 1e-4 * np.abs(ref).max()
    assert np.abs(h.cpu().numpy() - g["h"]).max() <= 1e-4 * max(1.0, np.abs(g["h"]).max())
    margin = np.abs(ref[:, 1] - ref[:, 0])
    assert not np.any((net.last_pred.cpu().numpy() != g["pred"]) & (margin > 1e-4))
