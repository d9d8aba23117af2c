"""evaluation counts, clustering scores and the tracking output.

CPU part: the oracle against the golden vectors of the reference's compute_P_R_F, and the two HOST entry points of the
library (EMI, text writer) against scikit-learn / numpy.  GPU part: the kernels through the C ABI against the oracle."""
import os

import numpy as np
import pytest
import torch

import gcn_mtmc_b200 as m
from oracle import eval_oracle as eo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_prf.npz")


def dev():
    return torch.device("cuda", 0)


# ------------------------------------------------------------------------------------------------------------ CPU
def test_prf_oracle_matches_reference_golden():
    g = np.load(GOLDEN)
    for tag in "abc":
        TP, FP, TN, FN, P, R, F, pc0, pc1 = eo.compute_P_R_F(torch.from_numpy(g[f"{tag}_preds"]), torch.from_numpy(g[f"{tag}_labels"]))
        assert [int(TP), int(FP), int(TN), int(FN)] == g[f"{tag}_counts"].tolist()
        got = np.array([float(P), float(R), float(F), float(pc0[0]), float(pc1[0])], dtype=np.float32)
        assert np.array_equal(got, g[f"{tag}_prf"])                       # same torch arithmetic: bit-exact


def test_expected_mutual_information_host_matches_sklearn():
    from sklearn.metrics.cluster import contingency_matrix, expected_mutual_information
    rng = np.random.default_rng(0)
    lib = m._lib.lib()
    for n, ka, kb in [(50, 3, 4), (400, 17, 9), (1000, 60, 75), (30, 1, 5), (7, 7, 7)]:
        a, b = rng.integers(0, ka, n), rng.integers(0, kb, n)
        cont = contingency_matrix(a, b, sparse=True)
        ref = expected_mutual_information(cont, n)
        ra = np.ascontiguousarray(np.ravel(cont.sum(axis=1)), dtype=np.int64)
        cb = np.ascontiguousarray(np.ravel(cont.sum(axis=0)), dtype=np.int64)
        got = lib.mpn_expected_mutual_information_host(ra.ctypes.data, ra.size, cb.ctypes.data, cb.size, n)
        assert abs(got - ref) <= 1e-12 * max(1.0, abs(ref)), (n, ka, kb, got, ref)


def test_save_mtmc_matches_numpy_savetxt(tmp_path):
    rng = np.random.default_rng(1)
    table = rng.integers(-5, 3000, size=(2500, 7)).astype(np.int64)
    table[0, 0] = 0
    table[1, 2] = -(2 ** 62)
    ours, ref = tmp_path / "a.txt", tmp_path / "b.txt"
    m.save_mtmc(ours, table)
    np.savetxt(ref, table, fmt="%d")                                       # main.py:114
    assert ours.read_bytes() == ref.read_bytes()
    tf = table[:50].astype(np.float64) + 0.75                              # '%d' truncates toward zero
    m.save_mtmc(ours, tf)
    np.savetxt(ref, tf, fmt="%d")
    assert ours.read_bytes() == ref.read_bytes()
    m.save_mtmc(ours, np.zeros((0, 7), dtype=np.int64))
    assert ours.read_bytes() == b""


def test_product_path_refuses_cpu_tensors():
    with pytest.raises(RuntimeError):
        m.compute_P_R_F(torch.zeros(4, dtype=torch.int64), torch.zeros(4))
    with pytest.raises(RuntimeError):
        m.evaluation.Contingency([0, 1], [1, 0], device="cpu")


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_compute_P_R_F_matches_golden_and_oracle():
    g = np.load(GOLDEN)
    for tag in "abc":
        preds, labels = torch.from_numpy(g[f"{tag}_preds"]), torch.from_numpy(g[f"{tag}_labels"])
        TP, FP, TN, FN, P, R, F, pc0, pc1 = m.compute_P_R_F(preds.to(dev()), labels.to(dev()))
        assert TP.is_cuda and P.is_cuda and TP.dtype == torch.int64
        assert [int(TP), int(FP), int(TN), int(FN)] == g[f"{tag}_counts"].tolist()
        got = np.array([float(P), float(R), float(F), float(pc0[0]), float(pc1[0])], dtype=np.float32)
        # counts are exact; the fp32 ratios follow the reference's own tensor arithmetic, which on CUDA divides a tensor by a
        # Python scalar as a multiplication by the reciprocal (1 ulp from the CPU run that produced the fixture)
        assert np.allclose(got, g[f"{tag}_prf"], rtol=3e-7, atol=0)
    gen = torch.Generator().manual_seed(5)
    E = 1_000_003
    labels = (torch.rand(E, generator=gen) < 0.2).long()
    preds = torch.where(torch.rand(E, generator=gen) < 0.9, labels, 1 - labels)
    labels[::1001] = 3                                                     # values other than 0/1 are ignored by every count
    preds[::777] = 2
    for pt, lt in [(torch.int64, torch.int64), (torch.uint8, torch.float32), (torch.float32, torch.uint8), (torch.int32, torch.float64)]:
        ref = eo.compute_P_R_F(preds.to(pt), labels.to(lt))
        got = m.compute_P_R_F(preds.to(pt).to(dev()), labels.to(lt).to(dev()))
        assert [int(v) for v in got[:4]] == [int(v) for v in ref[:4]]
        assert np.allclose([float(v) for v in got[4:7]] + [float(got[7][0]), float(got[8][0])],
                           [float(v) for v in ref[4:7]] + [float(ref[7][0]), float(ref[8][0])], rtol=3e-7, atol=0)
    # the reference's zero branches: no positive label / nothing predicted active / empty
    for preds, labels in [(torch.zeros(10).long(), torch.zeros(10)), (torch.zeros(10).long(), torch.ones(10)),
                          (torch.ones(10).long(), torch.zeros(10)), (torch.zeros(0).long(), torch.zeros(0))]:
        ref = eo.compute_P_R_F(preds, labels)
        got = m.compute_P_R_F(preds.to(dev()), labels.to(dev()))
        assert [int(v) for v in got[:4]] == [int(v) for v in ref[:4]]
        assert [float(v) for v in got[4:7]] == [float(v) for v in ref[4:7]]
        assert float(got[7][0]) == float(ref[7][0]) and float(got[8][0]) == float(ref[8][0])


@pytest.mark.gpu
@pytest.mark.parametrize("n,ka,kb,seed", [(300, 40, 37, 0), (5000, 700, 650, 1), (1000, 1, 1, 2), (1000, 1, 30, 3), (64, 64, 64, 4),
                                          (20_000, 3_000, 2_800, 5), (2, 2, 1, 6)])
def test_clustering_scores_match_sklearn(n, ka, kb, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, ka, n) * 3 - 7                                     # arbitrary (non-compact, negative) label values
    b = np.where(rng.random(n) < 0.8, (a + 7) // 3 % kb, rng.integers(0, kb, n)) + 100
    ref = eo.clustering_scores(a, b)
    got = m.clustering_scores(torch.from_numpy(a).to(dev()), torch.from_numpy(b))      # device and host inputs both accepted
    for k, v in ref.items():
        assert abs(got[k] - v) <= 1e-9 * max(1.0, abs(v)), (k, got[k], v)      # EMI: millions of lgamma terms, libm vs scipy
    assert abs(m.evaluation.adjusted_rand_score(a, b) - ref["adjusted_rand_score"]) <= 1e-12
    assert abs(m.evaluation.v_measure_score(a, b) - ref["v_measure_score"]) <= 1e-10
    c = m.evaluation.Contingency(a, b)
    from sklearn.metrics.cluster import contingency_matrix
    dense = contingency_matrix(a, b)
    r, cc = np.nonzero(dense)
    assert np.array_equal(c.rows, r) and np.array_equal(c.cols, cc) and np.array_equal(c.vals, dense[r, cc])      # bit-exact counts
    assert np.array_equal(c.a, dense.sum(1)) and np.array_equal(c.b, dense.sum(0))


@pytest.mark.gpu
def test_identical_and_empty_labelings():
    a = np.arange(50) // 5
    s = m.clustering_scores(a, a)
    assert all(abs(v - 1.0) <= 1e-12 for v in s.values())
    e = m.clustering_scores(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64))
    assert e["homogeneity_score"] == 1.0 and e["adjusted_rand_score"] == 1.0 and e["adjusted_mutual_info_score"] == 1.0


@pytest.mark.gpu
def test_tracking_table_matches_reference_loop(tmp_path):
    import pandas as pd
    rng = np.random.default_rng(9)
    n_cam, n_trk, M = 6, 400, 60_000
    node_cam = rng.integers(0, n_cam, n_trk)
    node_old = rng.integers(0, 150, n_trk)                                 # repeated (cam, id) keys: the last tracklet wins
    ID_pred = rng.integers(0, 90, n_trk)
    det = pd.DataFrame({"frame": rng.integers(0, 2000, M), "id": rng.integers(0, 170, M), "xmin": rng.integers(0, 1900, M),
                        "ymin": rng.integers(0, 1000, M), "width": rng.integers(5, 300, M), "height": rng.integers(5, 300, M),
                        "id_cam": rng.integers(0, n_cam + 1, M)})         # camera n_cam has no tracklet: ids stay
    # the reference, literally (inference.py:540-551)
    data_tracking = det.copy()
    for n in range(n_trk):
        data_tracking.loc[(det['id'] == node_old[n]).values & (det['id_cam'] == node_cam[n]).values, 'id'] = int(ID_pred[n])
    ref = data_tracking[['id_cam', 'id', 'frame', 'xmin', 'ymin', 'width', 'height']].values
    assert np.array_equal(eo.relabel_loop(det['id_cam'].values, det['id'].values, node_cam, node_old, ID_pred), ref[:, 1])
    got = m.tracking_table(det, node_cam, node_old, torch.from_numpy(ID_pred).to(dev()))
    assert got.dtype == np.int64 and np.array_equal(got, ref)
    ours, reff = tmp_path / "mtmc_a.txt", tmp_path / "mtmc_b.txt"
    m.save_mtmc(ours, got)
    np.savetxt(reff, ref, fmt='%d')                                        # main.py:114
    assert ours.read_bytes() == reff.read_bytes()
    with pytest.raises(m._lib.MpnError):                                   # ids the 20+44-bit join key cannot hold
        m.relabel_detections([0], [5], [0], [-1], [3], device=dev())
    assert m.relabel_detections([], [], [], [], [], device=dev()).numel() == 0
    assert m.relabel_detections([1, 2], [5, 6], [], [], [], device=dev()).tolist() == [5, 6]


# ------------------------------------------------------------------------------------------------------------ graph inputs (8f rows 1, 4)
def test_packed_reid_features_roundtrip_from_reference_pickles(tmp_path):
    """The reference's layout: ./reid_features/<scenario>/c<cam:03d>/<id:04d>/<file>_<model>.pkl, one pickled CPU tensor each."""
    import pickle
    gen = torch.Generator().manual_seed(3)
    cams = [1, 1, 1, 2, 2, 4]
    ids = [7, 12, 300, 7, 9, 12]
    feats = [torch.randn(96, generator=gen) for _ in cams]
    root = tmp_path / "reid_features"
    for c, t, f in zip(cams, ids, feats):
        p = m.graph_inputs.reid_feature_path(str(root), "S02", c, t, "bbox", "resnet101")
        os.makedirs(os.path.dirname(p))
        with open(p, "wb") as fo:
            pickle.dump(f, fo)                                             # libs/reid_feature_extraction.py:181-184
    packed = tmp_path / "S02.mpnfeat"
    m.pack_reid_features_from_pickles(str(packed), str(root), "S02", "bbox", "resnet101", cams, ids)
    x, cam, tid = m.read_packed_features(str(packed))
    assert cam.tolist() == cams and tid.tolist() == ids
    assert np.array_equal(np.asarray(x), torch.stack(feats).numpy())       # bit-exact
    m.pack_reid_features(str(packed), np.zeros((0, 5), dtype=np.float32), [], [])
    x0, c0, t0 = m.read_packed_features(str(packed))
    assert x0.shape == (0, 5) and c0.size == 0
    with pytest.raises(ValueError):
        m.pack_reid_features(str(packed), np.zeros((3, 5)), [1, 2], [1, 2, 3])
    with pytest.raises(RuntimeError):
        m.load_packed_features(str(packed), "cpu")


@pytest.mark.gpu
def test_load_packed_normalize_and_edge_labels(tmp_path):
    from oracle import mpn_oracle as mo
    x, ei, cam, ident = mo.synth_graph(90, 5, 3, D=64, planted=True)
    raw = x * (1.0 + torch.arange(64)) + 0.3                               # un-normalised features
    packed = tmp_path / "seq.mpnfeat"
    m.pack_reid_features(str(packed), raw, cam.numpy(), np.arange(90))
    xd, cam2, tid2 = m.load_packed_features(str(packed), dev())
    assert torch.equal(xd.cpu(), raw) and cam2.tolist() == cam.tolist()
    ref = torch.nn.functional.normalize(raw, p=2, dim=0)                   # inference.py:403-404
    got = m.normalize_columns(xd)
    assert (got.cpu() - ref).abs().max().item() <= 2e-7
    xn, _, _ = m.load_packed_features(str(packed), dev(), l2norm=True)
    assert torch.equal(xn, got)
    z = torch.zeros(5, 3, device=dev())
    assert torch.equal(m.normalize_columns(z), z)                          # eps: 0 / max(0, 1e-12) = 0
    # ground-truth edge labels, literally as inference.py:446-450 builds them
    labels = ident.numpy() if torch.is_tensor(ident) else np.asarray(ident)
    nodes = np.arange(90)
    e_np = ei.numpy()
    ref_l = np.asarray([1 if (labels[nodes == e_np[0][i]] == labels[nodes == e_np[1][i]]) else 0 for i in range(e_np.shape[1])],
                       dtype=np.float32)
    got_l = m.edge_labels(labels, ei.to(dev()))
    assert got_l.dtype == torch.float32 and np.array_equal(got_l.cpu().numpy(), ref_l)
    g = m.TrackletGraph.from_cameras(cam, dev())
    assert np.array_equal(m.edge_labels(labels, graph=g).cpu().numpy(), ref_l)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(1))
    assert np.array_equal(m.edge_labels(labels, ei[:, perm].to(dev())).cpu().numpy(), ref_l[perm.numpy()])      # caller's edge order
