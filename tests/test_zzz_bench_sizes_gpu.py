"""Parity at the sizes the benchmark runs (BASELINE.json configs[1], [2], [3]).  B200 only; named to run last (the heaviest file).

configs[1]: 4096 tracklets, 8 cameras, 2048-d features, E = 14,680,064 — tcgen05 edge features on 262,144 sampled edges and the
            whole one-call forward against the fp64 oracle evaluated on the device (bench.parity_gate: the gate bench.py itself
            runs before timing), and the fused probabilities bit-identical to torch.softmax.
configs[2]: 2000 S02-shaped graphs in one call, 50 of them against the single-graph path.
configs[3]: post-processing on 1,000,000 nodes against the plain-C oracle (bit-exact decisions and reference label integers) on
            distinct probabilities (no tie: all-clusters-per-round equals the reference's order exactly, so the C rounds oracle is
            a valid checker at this size), and under float32 ties through size-independent properties (the statement-order oracle
            needs hours at this size; it pins the same code path at 200 k nodes in tests/test_zz_c_oracle_gpu.py).
"""
import numpy as np
import pytest
import torch

import bench
from oracle import postproc_c as pc
from oracle import postproc_oracle as po
from tests._util import Data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import gcn_mtmc_b200 as mod
    mod._lib.require_device(0)
    return mod


def test_configs1_size_features_and_forward_vs_fp64_oracle(m):
    dev = torch.device("cuda", 0)
    net = bench.make_model(dev)
    x, ei = bench.device_graph(bench.NODES_1GPU, bench.CAMS, 0, dev)
    assert ei.shape[1] == 14680064
    res = bench.parity_gate(m, net, x, ei, dev)               # raises on failure
    assert res["checked"] and res["decisions_differ_outside_margin_band"] == 0 and res["prob1_bit_identical_to_torch_softmax"]
    assert res["logit_max_abs_err"] <= res["logit_tolerance"] and res["edge_feature_max_abs_err"] <= 1e-5
    # planted features (same-identity pairs: cancellation, refine list) at the same size
    from oracle import mpn_oracle as mo
    xp, eip, _, _ = mo.synth_graph(bench.NODES_1GPU, bench.CAMS, 1, planted=True)
    res = bench.parity_gate(m, net, xp.to(dev), eip.to(dev), dev)
    assert res["decisions_differ_outside_margin_band"] == 0


def test_configs2_size_batched_graphs(m):
    res = bench.extra_batched_graphs(m, torch.device("cuda", 0))          # raises when a checked graph differs
    assert res["max_rel_logit_diff_vs_single_graph_path_50_graphs"] <= 1e-4


def test_configs3_size_post_processing_vs_c_oracle(m):
    dev = torch.device("cuda", 0)
    n_nodes, cams = 1_000_000, 8
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, 7, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03,
                                                          single_dir=0.05)
    # distinct probabilities on the active edges: no tie, so the reference's order cannot matter
    rng = np.random.default_rng(1)
    a = pred > 0
    n_act = int(a.sum())
    bits = np.float32(0.55).view(np.uint32) + rng.permutation(n_act).astype(np.uint32)     # consecutive float32 values from 0.55 up
    prob_u = np.full(prob.size, 0.2, dtype=np.float32)
    prob_u[a] = bits.view(np.float32)
    assert np.unique(prob_u[a]).size == n_act and prob_u[a].max() < 1.0
    data = Data(x=torch.zeros(n_nodes, 1, device=dev), edge_index=torch.from_numpy(np.stack([src, dst])).to(dev))
    cfg = {"CUTTING": True, "PRUNING": True, "SPLITTING": True}
    lab_ref, act_ref = pc.post_processing(src, dst, pred, prob_u, cams, n_nodes, numbering="reference")
    ID, P = m.post_processing(cams, None, None, torch.from_numpy(pred).to(dev), None, dict(cfg), data, torch.from_numpy(prob_u).to(dev))
    st = m.split_stats()
    assert st["mode"] == "device_rounds" and st["tied_edges"] == 0
    assert np.array_equal(P.cpu().numpy(), act_ref) and np.array_equal(ID.numpy(), lab_ref)
    # float32 random probabilities: thousands of ties -> the reference's order on the host; properties that hold at any size
    ID, P = m.post_processing(cams, None, None, torch.from_numpy(pred).to(dev), None, dict(cfg), data, torch.from_numpy(prob).to(dev))
    st = m.split_stats()
    assert st["mode"] == "reference_order_host" and st["tied_edges"] > 0
    assert np.bincount(ID.numpy()).max() <= cams
    assert bool((P.cpu() <= torch.from_numpy(pred)).all())                          # only active edges were switched off
    ID2, P2 = m.post_processing(cams, None, None, P.clone(), None, {"CUTTING": False, "PRUNING": False, "SPLITTING": True}, data,
                                torch.from_numpy(prob).to(dev))
    assert torch.equal(P2, P) and np.array_equal(ID2.numpy(), ID.numpy())          # SPLITTING again: nothing oversized is left
    # CUT + PRUNE + CUT (order-free stages) against the C oracle on the tied probabilities too
    lab_c, act_c = pc.post_processing(src, dst, pred, prob, cams, n_nodes, splitting=False, numbering="reference")
    IDc, Pc = m.post_processing(cams, None, None, torch.from_numpy(pred).to(dev), None,
                                {"CUTTING": True, "PRUNING": True, "SPLITTING": False}, data, torch.from_numpy(prob).to(dev))
    assert np.array_equal(Pc.cpu().numpy(), act_c) and np.array_equal(IDc.numpy(), lab_c)
