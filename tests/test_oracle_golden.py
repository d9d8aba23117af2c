"""The oracle (oracle/) replayed against outputs of the real reference (tests/golden/*.npz).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import mpn_oracle as mo
from oracle import postproc_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MPN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "mpn_*.npz")))
POST_FILES = sorted(glob.glob(os.path.join(GOLDEN, "post_*.npz")))


from tests._util import load_mpn_case  # noqa: E402  (checks the regenerated inputs against the stored checksums)


def test_golden_present():
    assert len(MPN_FILES) >= 5 and len(POST_FILES) >= 6


@pytest.mark.parametrize("path", MPN_FILES, ids=[os.path.basename(p)[:-4] for p in MPN_FILES])
def test_mpn_oracle_matches_reference(path):
    g, params, sd, x, edge_index, _ = load_mpn_case(path)
    ea = mo.edge_features(x, edge_index)
    # fp32 tolerance: 2e-6 relative on the distance, 2e-6 absolute on 1-cos (values ~1)
    assert np.allclose(ea.numpy(), g["edge_attr"], rtol=2e-6, atol=2e-6)
    outs, h = mo.mpn_forward(sd, params, "resnet101", x, edge_index, torch.from_numpy(g["edge_attr"]))
    assert len(outs) == int(g["n_logits"][0])
    for i, o in enumerate(outs):
        ref = g[f"logits{i}"]
        tol = 1e-4 * np.abs(ref).max()              # north_star: logits within 1e-4 relative (of max |logit|)
        assert np.abs(o.numpy() - ref).max() <= tol
    assert np.abs(h.numpy() - g["h"]).max() <= 1e-4 * max(1.0, np.abs(g["h"]).max())
    prob, pred = mo.decide(outs[-1])
    margin = np.abs(g[f"logits{len(outs) - 1}"][:, 1] - g[f"logits{len(outs) - 1}"][:, 0])
    differ = pred.numpy() != g["pred"]
    assert not np.any(differ & (margin > 1e-4))
    # fp64 run of the oracle against the fp64 run of the reference: tight
    outs64, h64 = mo.mpn_forward(sd, params, "resnet101", x, edge_index, torch.from_numpy(g["edge_attr"]),
                                 dtype=torch.float64)
    for i, o in enumerate(outs64):
        assert np.abs(o.numpy() - g[f"logits64_{i}"]).max() <= 1e-9
    assert np.abs(h64.numpy() - g["h64"]).max() <= 1e-8


@pytest.mark.parametrize("path", POST_FILES, ids=[os.path.basename(p)[:-4] for p in POST_FILES])
def test_postproc_oracle_matches_reference(path):
    g = np.load(path)
    N, C, _seed = [int(v) for v in g["spec"]]
    src, dst = g["src"].astype(np.int64), g["dst"].astype(np.int64)
    prob, pred = g["prob1"], g["pred"].astype(np.int64)
    lab0, _ = po.scc_labels_reference(src, dst, pred, N)
    assert np.array_equal(lab0, g["labels_initial"])
    for tag, cfg in (("full", (True, True, True)), ("cut_only", (True, False, False)),
                     ("prune_only", (False, True, False)), ("split_only", (False, False, True)),
                     ("cut_prune", (True, True, False))):
        lab_s, act_s = po.post_processing_sequential(src, dst, pred, prob, C, N, *cfg)
        assert np.array_equal(act_s, g["pred_" + tag]), tag
        assert np.array_equal(lab_s, g["labels_" + tag]), tag
        lab_r, act_r = po.post_processing_rounds(src, dst, pred, prob, C, N, *cfg, numbering="reference")
        assert np.array_equal(act_r, g["pred_" + tag]), tag
        assert np.array_equal(lab_r, g["labels_" + tag]), tag
        lab_c, _ = po.post_processing_rounds(src, dst, pred, prob, C, N, *cfg, numbering="canonical")
        # canonical numbering: same partition
        assert len(set(zip(lab_c.tolist(), g["labels_" + tag].tolist()))) == len(set(lab_c.tolist()))


def test_tarjan_order_matches_networkx():
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(5, 40))
        m = int(rng.integers(1, 4 * n))
        src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
        keep = src != dst
        src, dst = src[keep], dst[keep]
        ours = po._tarjan_networkx_order(src, dst)
        theirs = list(nx.strongly_connected_components(nx.DiGraph(list(zip(src.tolist(), dst.tolist())))))
        assert [sorted(s) for s in ours] == [sorted(s) for s in theirs]


def test_rounds_equal_sequential_random():
    for seed in range(12):
        N, C = 30 + 5 * seed, 3 + seed % 3
        src, dst, prob, pred, _ = po.planted_prediction_graph(N, C, 1000 + seed, flip_on=0.06, flip_off=0.04,
                                                              single_dir=0.03, dense=True)
        lab_s, act_s = po.post_processing_sequential(src, dst, pred, prob, C, N)
        lab_r, act_r = po.post_processing_rounds(src, dst, pred, prob, C, N, numbering="reference")
        assert np.array_equal(act_s, act_r)
        assert np.array_equal(lab_s, lab_r)


def test_split_rounds_equal_sequential_unless_cross_cluster_tie():
    """Heavily tied probabilities (one decimal): the rounds formulation of SPLITTING (what the CUDA path runs) equals the
    reference's one-cluster-at-a-time loop in every run without a cross-cluster tie, and ``report_ties`` flags every run in
    which it does not.  CUT and PRUNE agree under ties as well (first index on equal probabilities)."""
    rng = np.random.default_rng(123)
    differ_flagged = ties_seen = 0
    for trial in range(500):
        n, C = int(rng.integers(6, 28)), int(rng.integers(2, 5))
        cam = np.sort(rng.integers(0, C, n))
        s, d = np.nonzero(cam[:, None] != cam[None, :])
        keep = rng.random(s.size) < rng.uniform(0.5, 1.0)
        s, d = s[keep], d[keep]
        if s.size == 0:
            continue
        prob = (np.round(rng.random(s.size) * 10) / 10).astype(np.float32)
        pred = (prob > 0.5).astype(np.int64)
        act = po.cut_sequential(s, d, pred)
        assert np.array_equal(act, po.cut_rounds(pred, po.reverse_edge_map(s, d, n)).astype(np.int64))
        pr_s = po.prune_sequential(s, d, act, prob, C, n)
        pr_r, changed = po.prune_rounds(s, d, act, prob, C, n)
        assert (pr_s is None) == (not changed)
        assert np.array_equal(act if pr_s is None else pr_s, pr_r.astype(np.int64))
        start = act if trial % 2 else pred
        sp_s = po.split_sequential(s, d, start, prob, C, n)
        sp_r, tie_rounds = po.split_rounds(s, d, start, prob, C, n, report_ties=True)
        ties_seen += tie_rounds > 0
        if not np.array_equal(sp_s, sp_r.astype(np.int64)):
            assert tie_rounds > 0, "rounds differ from the statement mirror without a cross-cluster tie (trial %d)" % trial
            differ_flagged += 1
    assert ties_seen > 0 and differ_flagged > 0        # the generator does reach the condition


def test_mpn_oracle_matches_reference_at_its_own_problem_size():
    """tests/golden/s02mpn_shipped_L1.npz: the unmodified reference (shipped L=1 configuration, 2048-d features) on one S02-shaped
    graph of 300 tracklets, 4 cameras, E = 67,500 directed edges (BASELINE configs[0]).  Same tolerances as the small cases."""
    path = os.path.join(GOLDEN, "s02mpn_shipped_L1.npz")
    g, params, sd, x, edge_index, _ = load_mpn_case(path)
    assert edge_index.shape[1] == 67500 and x.shape == (300, 2048)
    ea = mo.edge_features(x, edge_index)
    assert np.allclose(ea.numpy(), g["edge_attr"], rtol=2e-6, atol=2e-6)
    outs, h = mo.mpn_forward(sd, params, "resnet101", x, edge_index, torch.from_numpy(g["edge_attr"]))
    ref = g["logits0"]
    assert len(outs) == 1 and np.abs(outs[0].numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    assert np.abs(h.numpy() - g["h"]).max() <= 1e-4 * max(1.0, np.abs(g["h"]).max())
    _, pred = mo.decide(outs[-1])
    margin = np.abs(ref[:, 1] - ref[:, 0])
    assert not np.any((pred.numpy() != g["pred"]) & (margin > 1e-4))
