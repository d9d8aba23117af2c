"""The plain-C post-processing oracle (oracle/postproc_oracle.c) pinned on the CPU: against the outputs of the unmodified reference
(tests/golden/post_*.npz), against the Python restatement (statement mirror and rounds) on random predicted graphs, against
networkx's SCC emission order, and on the edge cases the reference meets (no edges, no active edges, one-directional cycles)."""
import glob
import os

import numpy as np
import pytest

from oracle import postproc_c as pc
from oracle import postproc_oracle as po

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
POST_FILES = sorted(glob.glob(os.path.join(GOLDEN, "post_*.npz")))
FLAG_SETS = (("full", (True, True, True)), ("cut_only", (True, False, False)), ("prune_only", (False, True, False)),
             ("split_only", (False, False, True)), ("cut_prune", (True, True, False)))


@pytest.mark.parametrize("path", POST_FILES, ids=[os.path.basename(p)[:-4] for p in POST_FILES])
def test_c_oracle_matches_reference_golden(path):
    g = np.load(path)
    N, C, _seed = [int(v) for v in g["spec"]]
    src, dst = g["src"].astype(np.int64), g["dst"].astype(np.int64)
    prob, pred = g["prob1"], g["pred"].astype(np.int64)
    lab0, n0 = pc.scc_labels(src, dst, pred, N)
    assert np.array_equal(lab0, g["labels_initial"]) and n0 == int(lab0.max()) + 1
    for tag, cfg in FLAG_SETS:
        lab, act = pc.post_processing(src, dst, pred, prob, C, N, *cfg, numbering="reference")
        assert np.array_equal(act, g["pred_" + tag]), tag                      # decisions: bit-exact
        assert np.array_equal(lab, g["labels_" + tag]), tag                    # label integers: bit-exact
        lab_c, act_c = pc.post_processing(src, dst, pred, prob, C, N, *cfg, numbering="canonical")
        assert np.array_equal(act_c, act)
        assert np.array_equal(lab_c, po.scc_partition_canonical(src, dst, act, N))


def test_c_oracle_equals_python_oracle_small_dense():
    for seed in range(12):
        N, C = 30 + 5 * seed, 3 + seed % 3
        src, dst, prob, pred, _ = po.planted_prediction_graph(N, C, 1000 + seed, flip_on=0.06, flip_off=0.04, single_dir=0.03,
                                                              dense=True)
        lab_s, act_s = po.post_processing_sequential(src, dst, pred, prob, C, N)
        lab_c, act_c = pc.post_processing(src, dst, pred, prob, C, N, numbering="reference")
        assert np.array_equal(act_s, act_c) and np.array_equal(lab_s, lab_c)


@pytest.mark.parametrize("n_nodes,seed", [(3000, 5), (20000, 6), (200000, 7)])
def test_c_oracle_equals_python_rounds_sparse(n_nodes, seed):
    C = 6
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, C, seed, flip_on=0.04, flip_off=0.03, single_dir=0.02)
    for cfg in ((True, True, True), (False, True, True), (True, False, True)):
        lab_r, act_r = po.post_processing_rounds(src, dst, pred, prob, C, n_nodes, *cfg, numbering="canonical")
        lab_c, act_c = pc.post_processing(src, dst, pred, prob, C, n_nodes, *cfg, numbering="canonical")
        assert np.array_equal(act_r, act_c) and np.array_equal(lab_r, lab_c)
    if n_nodes <= 20000:                                                        # (the Python Tarjan mirror is interpreter-bound)
        lab_r, act_r = po.post_processing_rounds(src, dst, pred, prob, C, n_nodes, numbering="reference")
        lab_c, act_c = pc.post_processing(src, dst, pred, prob, C, n_nodes, numbering="reference")
        assert np.array_equal(act_r, act_c) and np.array_equal(lab_r, lab_c)


def test_c_oracle_stages_equal_python_stages():
    N, C = 4000, 5
    src, dst, prob, pred, _ = po.planted_prediction_graph(N, C, 11, flip_on=0.05, flip_off=0.03, single_dir=0.03)
    rev = po.reverse_edge_map(src, dst, N)
    assert np.array_equal(pc.reverse_edge_map(src, dst, N), rev)
    cut_py = po.cut_rounds(pred, rev)
    assert np.array_equal(pc.cut(src, dst, pred, N), cut_py.astype(np.int64))
    pr_py, ch_py = po.prune_rounds(src, dst, cut_py, prob, C, N)
    pr_c, ch_c = pc.prune(src, dst, cut_py, prob, C, N)
    assert ch_c == ch_py and np.array_equal(pr_c, pr_py.astype(np.int64))
    sp_py = po.split_rounds(src, dst, pr_py, prob, C, N)
    assert np.array_equal(pc.split(src, dst, pr_py, prob, C, N), sp_py.astype(np.int64))
    # unsorted edge order (the reverse map then sorts its keys)
    perm = np.random.default_rng(0).permutation(src.size)
    rev_p = pc.reverse_edge_map(src[perm], dst[perm], N)
    assert np.array_equal(rev_p, po.reverse_edge_map(src[perm], dst[perm], N))


def test_c_tarjan_order_matches_networkx():
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(5, 40))
        m = int(rng.integers(1, 4 * n))
        src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
        keep = src != dst
        src, dst = src[keep], dst[keep]
        if src.size == 0:
            continue
        G = nx.DiGraph(list(zip(src.tolist(), dst.tolist())))
        sccs = sorted(nx.strongly_connected_components(G), key=len)            # utils.py:31
        expect = np.full(n, -1, dtype=np.int64)
        for k, s in enumerate(sccs):
            expect[list(s)] = k
        k = len(sccs)
        for i in range(n):                                                      # utils.py:34-42
            if expect[i] < 0:
                expect[i] = k
                k += 1
        lab, n_comp = pc.scc_labels(src, dst, np.ones(src.size, dtype=np.int64), n)
        assert n_comp == k and np.array_equal(lab, expect)


def test_c_oracle_edge_cases():
    e = np.zeros(0, dtype=np.int64)
    lab, act = pc.post_processing(e, e, e, np.zeros(0, dtype=np.float32), 4, 5, numbering="reference")
    assert act.size == 0 and np.array_equal(lab, np.arange(5))
    src, dst = np.array([0, 1, 2, 3]), np.array([1, 2, 0, 0])                  # a one-directional 3-cycle + a tail
    prob = np.array([0.9, 0.8, 0.7, 0.6], dtype=np.float32)
    lab, act = pc.post_processing(src, dst, np.zeros(4, dtype=np.int64), prob, 4, 4, numbering="reference")
    assert not act.any() and np.array_equal(lab, np.arange(4))                  # nothing active: singletons in index order
    lab, _ = pc.scc_labels(src, dst, np.ones(4, dtype=np.int64), 4)
    lab_py, _ = po.scc_labels_reference(src, dst, np.ones(4, dtype=np.int64), 4)
    assert np.array_equal(lab, lab_py)
    lab, act = pc.post_processing(src, dst, np.ones(4, dtype=np.int64), prob, 4, 4, cutting=True, pruning=False,
                                  splitting=False, numbering="reference")
    assert not act.any()                                                         # no edge has its reverse: CUT removes all
    with pytest.raises(RuntimeError):
        pc.post_processing(np.array([0, 9]), np.array([1, 0]), np.ones(2, dtype=np.int64), prob[:2], 4, 4)   # node id out of range


def test_c_oracle_pins_the_gpu_case():
    """The input of tests/test_zz_c_oracle_gpu.py (GPU vs C oracle): here the C oracle against the Python rounds oracle."""
    from tests.test_zz_c_oracle_gpu import GPU_CASE as c
    src, dst, prob, pred, _ = po.planted_prediction_graph(c["n_nodes"], c["cams"], c["seed"], n_extra_per_node=c["n_extra_per_node"],
                                                          flip_on=c["flip_on"], flip_off=c["flip_off"], single_dir=c["single_dir"])
    lab_r, act_r = po.post_processing_rounds(src, dst, pred, prob, c["cams"], c["n_nodes"], numbering="reference")
    lab_c, act_c = pc.post_processing(src, dst, pred, prob, c["cams"], c["n_nodes"], numbering="reference")
    assert np.array_equal(act_r, act_c) and np.array_equal(lab_r, lab_c)
    assert np.bincount(lab_c).max() <= c["cams"] and int(act_c.sum()) > 0


def test_c_split_sequential_is_the_statement_mirror_under_ties():
    """po_split_sequential (the reference's own order, one cluster at a time) against the Python statement mirror on heavily
    tied probabilities, where the order matters."""
    rng = np.random.default_rng(5)
    n_checked = 0
    for trial in range(300):
        n, C = int(rng.integers(6, 28)), int(rng.integers(2, 5))
        cam = np.sort(rng.integers(0, C, n))
        s, d = np.nonzero(cam[:, None] != cam[None, :])
        keep = rng.random(s.size) < rng.uniform(0.5, 1.0)
        s, d = s[keep], d[keep]
        if s.size == 0:
            continue
        prob = (np.round(rng.random(s.size) * 10) / 10).astype(np.float32)
        pred = (prob > 0.5).astype(np.int64)
        start = po.cut_sequential(s, d, pred) if trial % 2 else pred
        assert np.array_equal(pc.split_sequential(s, d, start, prob, C, n), po.split_sequential(s, d, start, prob, C, n))
        n_checked += 1
    assert n_checked > 250


@pytest.mark.parametrize("n_nodes,cams,seed", [(3000, 6, 1), (20000, 8, 2), (20000, 8, 5)])
def test_rounds_equal_the_reference_order_on_the_gpu_test_graphs(n_nodes, cams, seed):
    """The graphs of the GPU parity tests (tests/test_gpu_parity.py, same generator arguments): the rounds formulation the CUDA
    path implements gives exactly what the reference's one-cluster-at-a-time SPLITTING gives (the exact order is affordable in
    C at this size), including the graph on which split_rounds reports cross-cluster ties."""
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, seed, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03,
                                                          single_dir=0.05)
    if n_nodes == 20000 and seed == 5:                                          # (the sharded test sorts its edges)
        order = np.lexsort((dst, src))
        src, dst, prob, pred = src[order], dst[order], prob[order], pred[order]
    act = pc.cut(src, dst, pred, n_nodes)
    act, _ = pc.prune(src, dst, act, prob, cams, n_nodes)
    act = pc.cut(src, dst, act, n_nodes)
    exact = pc.split_sequential(src, dst, act, prob, cams, n_nodes)
    rounds = pc.split(src, dst, act, prob, cams, n_nodes)
    assert np.array_equal(exact, rounds)
    lab, act_full = pc.post_processing(src, dst, pred, prob, cams, n_nodes, numbering="reference")
    assert np.array_equal(act_full, exact)


def test_c_split_hybrid_is_exact_under_ties():
    """po_split_hybrid — the reference's order only where a probability tie can make the order matter, all clusters per iteration
    elsewhere — equals the reference-order SPLITTING on tied graphs on which the plain rounds formulation does not."""
    n_checked = rounds_differ = 0
    for decimals in (1, 2, 3):
        rng = np.random.default_rng(7 + decimals)
        for trial in range(250):
            n, C = int(rng.integers(8, 40)), int(rng.integers(2, 5))
            cam = np.sort(rng.integers(0, C, n))
            s, d = np.nonzero(cam[:, None] != cam[None, :])
            keep = rng.random(s.size) < rng.uniform(0.4, 1.0)
            s, d = s[keep], d[keep]
            if s.size == 0:
                continue
            q = 10 ** decimals
            prob = (np.round(rng.random(s.size) * q) / q).astype(np.float32)
            pred = (prob > 0.5).astype(np.int64)
            start = po.cut_sequential(s, d, pred) if trial % 2 else pred
            exact = pc.split_sequential(s, d, start, prob, C, n)
            hybrid, stats = pc.split_hybrid(s, d, start, prob, C, n)
            assert np.array_equal(hybrid, exact), (decimals, trial, stats)
            rounds_differ += not np.array_equal(pc.split(s, d, start, prob, C, n), exact)
            n_checked += 1
    assert n_checked > 600 and rounds_differ > 0


def test_c_split_hybrid_on_a_20k_node_graph():
    n_nodes, cams = 20000, 8
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, 5, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03,
                                                          single_dir=0.05)
    act = pc.cut(src, dst, pred, n_nodes)
    act, _ = pc.prune(src, dst, act, prob, cams, n_nodes)
    act = pc.cut(src, dst, act, n_nodes)
    hybrid, stats = pc.split_hybrid(src, dst, act, prob, cams, n_nodes)
    assert stats["tie_values"] > 0 and stats["tainted_steps"] > 0
    assert np.array_equal(hybrid, pc.split_sequential(src, dst, act, prob, cams, n_nodes))


def _ties_cases():
    g = np.load(os.path.join(GOLDEN, "ties_cases.npz"))
    for i in range(int(g["n_cases"][0])):
        N, C = [int(v) for v in g[f"k{i}_spec"]]
        yield (i, g[f"k{i}_src"].astype(np.int64), g[f"k{i}_dst"].astype(np.int64), g[f"k{i}_prob1"], g[f"k{i}_pred"].astype(np.int64),
               C, N, {k[len(f"k{i}_"):]: g[k] for k in g.files if k.startswith(f"k{i}_labels_") or k.startswith(f"k{i}_pred_")})


def test_reference_order_under_ties_matches_the_reference_itself():
    """tests/golden/ties_cases.npz: outputs of the UNMODIFIED reference on graphs with exact probability ties (two decimals), half of
    them graphs on which the rounds formulation gives other decisions.  The statement mirror, the C restatement of the reference's
    order, the hybrid and the product's host SPLITTING must all reproduce the reference bit for bit; the rounds formulation is
    checked to differ on at least one case (that is what these fixtures are for)."""
    import ctypes as C_
    import gcn_mtmc_b200 as m
    lib = m._lib.lib()
    rounds_differ = n = 0
    for i, src, dst, prob, pred, C, N, ref in _ties_cases():
        for tag, cfg in (("full", (True, True, True)), ("split_only", (False, False, True)), ("prune_split", (False, True, True))):
            lab_s, act_s = po.post_processing_sequential(src, dst, pred, prob, C, N, *cfg)           # Python statement mirror
            assert np.array_equal(act_s, ref["pred_" + tag]) and np.array_equal(lab_s, ref["labels_" + tag]), (i, tag)
            # C: CUT / PRUNE / CUT as rounds (equal to the statements, ties included), SPLITTING in the reference's order
            act = pred
            if cfg[0]:
                act = pc.cut(src, dst, act, N)
            if cfg[1]:
                act, _ = pc.prune(src, dst, act, prob, C, N)
            if cfg[0]:
                act = pc.cut(src, dst, act, N)
            exact = pc.split_sequential(src, dst, act, prob, C, N)
            assert np.array_equal(exact, ref["pred_" + tag]), (i, tag)
            assert np.array_equal(pc.scc_labels(src, dst, exact, N)[0], ref["labels_" + tag]), (i, tag)
            assert np.array_equal(pc.split_hybrid(src, dst, act, prob, C, N)[0], exact), (i, tag)
            a = np.flatnonzero(act)                                                                  # product: host SPLITTING
            if a.size:
                s32, d32 = np.ascontiguousarray(src[a], dtype=np.int32), np.ascontiguousarray(dst[a], dtype=np.int32)
                p32 = np.ascontiguousarray(prob[a], dtype=np.float32)
                keep = np.empty(a.size, dtype=np.uint8)
                m._lib.check(lib.mpn_split_exact_host(s32.ctypes.data, d32.ctypes.data, p32.ctypes.data, a.size, N, C,
                                                      keep.ctypes.data, None))
                got = act.copy()
                got[a[keep == 0]] = 0
                assert np.array_equal(got, ref["pred_" + tag]), (i, tag)
            rounds_differ += not np.array_equal(pc.split(src, dst, act, prob, C, N), exact)
            n += 1
    assert n == 30 and rounds_differ >= 5


S02_FLAG_SETS = FLAG_SETS


def load_s02_case():
    g = np.load(os.path.join(GOLDEN, "s02post_n300_c4.npz"))
    N, C, _seed = [int(v) for v in g["spec"]]
    return g, N, C, g["src"].astype(np.int64), g["dst"].astype(np.int64), g["prob1"], g["pred"].astype(np.int64)


def test_oracles_match_the_reference_at_its_own_problem_size():
    """tests/golden/s02post_n300_c4.npz: the unmodified reference's post_processing on one S02-shaped predicted graph (BASELINE
    configs[0]: 300 tracklets, 4 cameras, E = 67,500 directed edges, 3,896 active; 6.6 minutes of reference time for the five flag
    sets).  The C oracle (rounds and reference order), the numpy rounds and the statement mirror give its decisions and label
    integers bit for bit."""
    g, N, C, src, dst, prob, pred = load_s02_case()
    assert src.size == 67500 and N == 300 and C == 4
    assert np.array_equal(pc.scc_labels(src, dst, pred, N)[0], g["labels_initial"])
    for tag, cfg in S02_FLAG_SETS:
        lab, act = pc.post_processing(src, dst, pred, prob, C, N, *cfg, numbering="reference")
        assert np.array_equal(act, g["pred_" + tag]) and np.array_equal(lab, g["labels_" + tag]), tag
        lab, act = po.post_processing_rounds(src, dst, pred, prob, C, N, *cfg, numbering="reference")
        assert np.array_equal(act, g["pred_" + tag]) and np.array_equal(lab, g["labels_" + tag]), tag
    lab, act = po.post_processing_sequential(src, dst, pred, prob, C, N)
    assert np.array_equal(act, g["pred_full"]) and np.array_equal(lab, g["labels_full"])
    a = pc.cut(src, dst, pc.prune(src, dst, pc.cut(src, dst, pred, N), prob, C, N)[0], N)
    assert np.array_equal(pc.split_sequential(src, dst, a, prob, C, N), g["pred_full"])
