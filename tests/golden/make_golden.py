"""Generate golden vectors by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py        # needs /root/reference; writes tests/golden/*.npz

The reference ships no tests or fixtures, so these files are what pins the oracle
(and through it the CUDA path).  Inputs are regenerated from seeds by ``oracle/`` generators; each fixture
also stores input checksums so drift in the generators is detected, and small inputs are stored verbatim.
"""
import contextlib
import copy
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shims                                   # noqa: E402
from oracle import mpn_oracle as mo                # noqa: E402
from oracle import postproc_oracle as po           # noqa: E402

MPN_CASES = {
    # name: (N, C, graph seed, weight seed, L, n_cls, node_in_dim, node_fc_dims, planted, affine_jitter)
    "mpn_shipped_L1": (48, 4, 11, 1, 1, 1, 2048, (1024, 512, 128), False, True),
    "mpn_shipped_L1_planted": (60, 4, 12, 2, 1, 1, 2048, (1024, 512, 128), True, False),
    "mpn_small_L4_cls2": (40, 4, 13, 3, 4, 2, 64, (48, 40), True, True),
    "mpn_small_L0": (24, 3, 14, 4, 0, 1, 64, (48,), False, True),
    "mpn_small_L3_cls3_c5": (55, 5, 15, 5, 3, 3, 96, (64, 48, 40), True, True),
    # node_agg_fn variants (models/mpn.py:193-202), on a thinned graph with rows that lost all their edges:
    # (..., agg, thin)
    "mpn_small_L2_mean": (44, 4, 16, 6, 2, 2, 64, (48, 40), True, True, "mean", True),
    "mpn_small_L3_max": (44, 4, 17, 7, 3, 1, 64, (48, 40), True, True, "max", True),
    "mpn_shipped_L1_max": (40, 4, 18, 8, 1, 1, 2048, (1024, 512, 128), False, True, "max", False),
    # reattach_initial_nodes / reattach_initial_edges (models/mpn.py:207-215, 283-287): (..., agg, thin, re_n, re_e)
    "mpn_small_L3_reattach_both": (44, 4, 19, 9, 3, 2, 64, (48, 40), True, True, "sum", False, True, True),
    "mpn_small_L2_reattach_nodes": (40, 4, 20, 10, 2, 1, 64, (48, 40), True, True, "sum", True, True, False),
    "mpn_small_L3_reattach_edges": (40, 4, 21, 11, 3, 3, 64, (48, 40), True, True, "mean", False, False, True),
    "mpn_shipped_L1_reattach_both": (36, 4, 22, 12, 1, 1, 2048, (1024, 512, 128), False, True, "sum", False, True, True),
    # the reference's own problem size (BASELINE configs[0]): one S02-shaped graph, 300 tracklets, 4 cameras, 2048-d features,
    # E = 67,500, the shipped L=1 model.  Not named mpn_*: the parametrised tests that glob mpn_*.npz stay as they were measured
    # on the GPU; this fixture has its own tests (tests/test_oracle_golden.py, tests/test_zz_c_oracle_gpu.py)
    "s02mpn_shipped_L1": (300, 4, 23, 13, 1, 1, 2048, (1024, 512, 128), True, True),
}
AGG_CODE = {"sum": 0, "mean": 1, "max": 2}

POST_CASES = {
    # name: (N, C, seed, flip_on, flip_off, single_dir)
    "post_n40_c4": (40, 4, 101, 0.04, 0.04, 0.02),
    "post_n64_c4": (64, 4, 102, 0.06, 0.03, 0.03),
    "post_n90_c5": (90, 5, 103, 0.05, 0.05, 0.02),
    "post_n120_c4_noisy": (120, 4, 104, 0.10, 0.05, 0.05),
    "post_n36_c3": (36, 3, 105, 0.08, 0.02, 0.04),
    "post_n50_c4_clean": (50, 4, 106, 0.0, 0.0, 0.0),
    # the reference's own problem size (BASELINE configs[0]: one S02-shaped graph, 300 tracklets, 4 cameras, E = 67,500); 6.6 minutes
    # of reference time.  Not named post_*: the parametrised tests that glob post_*.npz stay as they were measured on the GPU, this
    # one has its own tests (tests/test_c_oracle.py, tests/test_zz_c_oracle_gpu.py)
    "s02post_n300_c4": (300, 4, 107, 0.03, 0.03, 0.02),
}


def run_mpn_case(MOTMPNet, name, spec):
    N, C, gseed, wseed, L, n_cls, din, fcd, planted, jitter = spec[:10]
    agg, thin = (spec[10], spec[11]) if len(spec) > 10 else ("sum", False)
    re_n, re_e = (spec[12], spec[13]) if len(spec) > 12 else (False, False)
    params = mo.shipped_model_params(L, n_cls, din, fcd)
    params["node_agg_fn"] = agg
    params["reattach_initial_nodes"], params["reattach_initial_edges"] = re_n, re_e
    x, edge_index, cam, ident = mo.synth_graph(N, C, gseed, D=din, planted=planted)
    if thin:
        edge_index = mo.thin_edges(edge_index, gseed)
    sd = mo.init_weights(params, "resnet101", wseed, affine_jitter=jitter)
    torch.manual_seed(0)
    model = MOTMPNet(copy.deepcopy(params), None, "resnet101").eval()
    assert list(model.state_dict().keys()) == list(sd.keys()), "oracle key layout differs from the reference"
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    with torch.no_grad():
        # edge features exactly as inference.py:453-456
        F = torch.nn.functional
        d = F.pairwise_distance(x[edge_index[0]], x[edge_index[1]]).view(-1, 1)
        c = 1 - F.cosine_similarity(x[edge_index[0]], x[edge_index[1]]).view(-1, 1)
        edge_attr = torch.cat((d, c), dim=1)
        Data = sys.modules["torch_geometric.data"].Data
        data = Data(x=x, edge_index=edge_index, edge_attr=edge_attr)
        out, h = model(data)
        logits = [t.numpy() for t in out["classified_edges"]]
        prob = torch.softmax(out["classified_edges"][-1], dim=1).numpy()       # inference.py:475-479
        pred = torch.argmax(out["classified_edges"][-1], dim=1).numpy()
        m64 = copy.deepcopy(model).double()
        out64, h64 = m64(Data(x=x.double(), edge_index=edge_index, edge_attr=edge_attr.double()))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        spec=np.array([N, C, gseed, wseed, L, n_cls, din, int(planted), int(jitter), AGG_CODE[agg], int(thin), int(re_n), int(re_e)],
                      dtype=np.int64),
        fc_dims=np.array(fcd, dtype=np.int64),
        x_checksum=np.array([x.double().sum().item(), x.double().abs().sum().item()]),
        w_checksum=np.array([sum(v.double().sum().item() for v in sd.values())]),
        edge_index=edge_index.numpy().astype(np.int32),
        edge_attr=edge_attr.numpy(),
        h=h.numpy(), prob=prob, pred=pred,
        n_logits=np.array([len(logits)]),
        **{f"logits{i}": l for i, l in enumerate(logits)},
        **{f"logits64_{i}": t.numpy() for i, t in enumerate(out64["classified_edges"])},
        h64=h64.numpy(),
    )
    if name.startswith("s02"):                 # the large fixture keeps the fp32 outputs only (1.2 MB instead of 2.4)
        path = os.path.join(HERE, name + ".npz")
        g = np.load(path)
        keep = {k: g[k] for k in g.files if not k.startswith("logits64_") and k not in ("h64", "prob")}
        keep["pred"] = keep["pred"].astype(np.int8)
        np.savez_compressed(path, **keep)
    print(name, "E=%d" % edge_index.shape[1], "max|logit|=%.3f" % np.abs(logits[-1]).max(),
          "active=%d" % int(pred.sum()))


def run_post_case(ref_utils, ref_inference, name, spec):
    import networkx as nx
    N, C, seed, f_on, f_off, sdir = spec
    src, dst, prob1, pred, cam = po.planted_prediction_graph(N, C, seed, flip_on=f_on, flip_off=f_off,
                                                             single_dir=sdir, dense=True)
    Data = sys.modules["torch_geometric.data"].Data
    edge_index = torch.from_numpy(np.stack([src, dst]))
    data = Data(x=torch.zeros(N, 1), edge_index=edge_index)
    edge_list = edge_index.numpy()
    preds_prob = torch.from_numpy(np.stack([1 - prob1, prob1], axis=1))
    out = {}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        predictions = torch.from_numpy(pred.copy())
        act_list = [(edge_list[0][p], edge_list[1][p]) for p in torch.where(predictions == 1)[0]]
        ID0, _ = ref_utils.compute_SCC_and_Clusters(nx.DiGraph(act_list), N)
        out["labels_initial"] = ID0.numpy()
        for tag, cfg in (("full", (True, True, True)), ("cut_only", (True, False, False)),
                         ("prune_only", (False, True, False)), ("split_only", (False, False, True)),
                         ("cut_prune", (True, True, False))):
            predictions = torch.from_numpy(pred.copy())
            act_list = [(edge_list[0][p], edge_list[1][p]) for p in torch.where(predictions == 1)[0]]
            CONFIG = {"CUTTING": cfg[0], "PRUNING": cfg[1], "SPLITTING": cfg[2]}
            ID, P = ref_inference.post_processing(C, ID0.clone(), act_list, predictions, edge_list, CONFIG, data, preds_prob)
            out["labels_" + tag] = ID.numpy()
            out["pred_" + tag] = P.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"),
                        spec=np.array([N, C, seed], dtype=np.int64), src=src.astype(np.int32), dst=dst.astype(np.int32),
                        prob1=prob1, pred=pred.astype(np.int8), **out)
    print(name, "E=%d active %d -> %d, clusters %d" % (src.size, pred.sum(), out["pred_full"].sum(),
                                                       out["labels_full"].max() + 1))


def tied_graphs(n_cases=10, seed=9):
    """Small cross-camera graphs with two-decimal probabilities: exact ties between edges of different clusters, where the ORDER
    in which the reference's splitting visits clusters decides the result.  Half of the cases are chosen so that the
    all-clusters-per-round formulation gives different decisions than the statement mirror (search on the CPU, deterministic)."""
    rng = np.random.default_rng(seed)
    differ, same = [], []
    for _ in range(4000):
        n, C = int(rng.integers(8, 40)), int(rng.integers(2, 5))
        cam = np.sort(rng.integers(0, C, n))
        s, d = np.nonzero(cam[:, None] != cam[None, :])
        keep = rng.random(s.size) < rng.uniform(0.4, 1.0)
        s, d = s[keep], d[keep]
        if s.size == 0:
            continue
        prob = (np.round(rng.random(s.size) * 100) / 100).astype(np.float32)
        pred = (prob > 0.5).astype(np.int64)
        seq = po.split_sequential(s, d, pred, prob, C, n)
        rnd = po.split_rounds(s, d, pred, prob, C, n).astype(np.int64)
        bucket = same if np.array_equal(seq, rnd) else differ
        if len(bucket) < n_cases // 2:
            bucket.append((s.astype(np.int64), d.astype(np.int64), prob, pred, C, n))
        if len(differ) >= n_cases // 2 and len(same) >= n_cases // 2:
            break
    return differ + same


def run_ties_cases(ref_utils, ref_inference):
    """The unmodified reference on the tied graphs: pins the behaviour under probability ties (utils.py:96-98,112)."""
    import networkx as nx
    Data = sys.modules["torch_geometric.data"].Data
    out = {}
    cases = tied_graphs()
    sink = io.StringIO()
    for i, (src, dst, prob1, pred, C, N) in enumerate(cases):
        edge_index = torch.from_numpy(np.stack([src, dst]))
        data = Data(x=torch.zeros(N, 1), edge_index=edge_index)
        edge_list = edge_index.numpy()
        preds_prob = torch.from_numpy(np.stack([1 - prob1, prob1], axis=1))
        out.update({f"k{i}_spec": np.array([N, C], dtype=np.int64), f"k{i}_src": src.astype(np.int32), f"k{i}_dst": dst.astype(np.int32),
                    f"k{i}_prob1": prob1, f"k{i}_pred": pred.astype(np.int8)})
        with contextlib.redirect_stdout(sink):
            for tag, cfg in (("full", (True, True, True)), ("split_only", (False, False, True)), ("prune_split", (False, True, True))):
                predictions = torch.from_numpy(pred.copy())
                act_list = [(edge_list[0][p], edge_list[1][p]) for p in torch.where(predictions == 1)[0]]
                ID0, _ = ref_utils.compute_SCC_and_Clusters(nx.DiGraph(act_list), N)
                CONFIG = {"CUTTING": cfg[0], "PRUNING": cfg[1], "SPLITTING": cfg[2]}
                ID, P = ref_inference.post_processing(C, ID0.clone(), act_list, predictions, edge_list, CONFIG, data, preds_prob)
                out[f"k{i}_labels_{tag}"] = ID.numpy()
                out[f"k{i}_pred_{tag}"] = P.numpy().astype(np.int8)
    out["n_cases"] = np.array([len(cases)], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "ties_cases.npz"), **out)
    print("ties_cases", len(cases), "graphs")


def run_eval_case(ref_inference):
    """compute_P_R_F of the reference (inference.py:20-66) on CPU tensors; every count is non-zero so that none of its
    `.cuda()` zero constants is reached."""
    g = torch.Generator().manual_seed(77)
    cases = {}
    for tag, E, p1 in (("a", 5000, 0.3), ("b", 257, 0.6), ("c", 64, 0.5)):
        labels = (torch.rand(E, generator=g) < p1).float()                 # labels_edges_GT is a float tensor (dataset.py)
        preds = torch.where(torch.rand(E, generator=g) < 0.85, labels, 1 - labels).long()
        TP, FP, TN, FN, P, R, F, pc0, pc1 = ref_inference.compute_P_R_F(preds, labels)
        cases.update({f"{tag}_preds": preds.numpy(), f"{tag}_labels": labels.numpy(),
                      f"{tag}_counts": np.array([int(TP), int(FP), int(TN), int(FN)], dtype=np.int64),
                      f"{tag}_prf": np.array([float(P), float(R), float(F), float(pc0[0]), float(pc1[0])], dtype=np.float32)})
    np.savez_compressed(os.path.join(HERE, "eval_prf.npz"), **cases)
    print("eval_prf", {k: v.tolist() for k, v in cases.items() if k.endswith("counts")})


def main():
    if not ref_shims.reference_available():
        raise SystemExit("reference not found at %s" % ref_shims.REFERENCE_ROOT)
    MOTMPNet, ref_utils, ref_inference, _ = ref_shims.load_reference()
    only = set(sys.argv[1:])
    for name, spec in MPN_CASES.items():
        if not only or name in only:
            run_mpn_case(MOTMPNet, name, spec)
    for name, spec in POST_CASES.items():
        if not only or name in only:
            run_post_case(ref_utils, ref_inference, name, spec)
    if not only or "eval_prf" in only:
        run_eval_case(ref_inference)
    if not only or "ties_cases" in only:
        run_ties_cases(ref_utils, ref_inference)


if __name__ == "__main__":
    main()
