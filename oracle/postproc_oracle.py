"""CPU oracle for CUTTING / PRUNING / SPLITTING + SCC labelling.  TEST INFRASTRUCTURE ONLY.

Two restatements of the reference post-processing (integer / index work, bit-exact bar):

* ``*_sequential`` functions follow the reference statement by statement (same loop order,
  same snapshot semantics, same label numbering); quadratic, for small graphs only.
* ``*_rounds`` functions are the vectorised "parallel rounds" formulation
  that scales to 1e8 edges; ``tests/test_oracle_golden.py (and tests/test_c_oracle.py for the C restatement)`` checks both against each other and
  against golden outputs of the real reference (``tests/golden/make_golden.py``).

Under exact probability ties across oversized clusters the rounds formulation of SPLITTING can differ from the reference's
one-cluster-at-a-time order; see ``split_rounds`` (``report_ties``) and ``oracle/postproc_oracle.c::po_split_sequential``.

Parity status: pinned only by outputs of the reference itself generated in the build container
("parity unpinned" by upstream tests: there are none).

Reference lines followed:
  * ``compute_SCC_and_Clusters``  utils.py:30-52  (networkx SCC order, ``sorted(key=len)``, isolated nodes last)
  * ``remove_edges_single_direction``  utils.py:125-142
  * ``pruning``  utils.py:144-339 (live lines 161-188, 277-317)
  * ``splitting``  utils.py:54-123
  * ``post_processing``  inference.py:70-169
  * networkx ``strongly_connected_components`` (networkx 2.5.1 pinned in env_gnn.yml:76; non-recursive
    Tarjan/Nuutila, source not vendored): restated in ``_tarjan_networkx_order``.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# SCC with networkx's emission order
# --------------------------------------------------------------------------------------
def _tarjan_networkx_order(src: np.ndarray, dst: np.ndarray):
    """SCCs of DiGraph(list(zip(src,dst))) in the order nx.strongly_connected_components yields them.

    Node iteration order = first appearance scanning (u, v) per edge in list order; successor
    order = order of first insertion of (u, v).  Returns list of lists of node ids.
    """
    order, adj = [], {}
    for u, v in zip(src.tolist(), dst.tolist()):
        if u not in adj:
            adj[u] = {}
            order.append(u)
        if v not in adj:
            adj[v] = {}
            order.append(v)
        adj[u][v] = None
    preorder, lowlink, found = {}, {}, set()
    scc_queue, out = [], []
    i = 0
    nbr_iter = {v: iter(adj[v]) for v in order}
    for source in order:
        if source in found:
            continue
        queue = [source]
        while queue:
            v = queue[-1]
            if v not in preorder:
                i += 1
                preorder[v] = i
            done = True
            for w in nbr_iter[v]:
                if w not in preorder:
                    queue.append(w)
                    done = False
                    break
            if done:
                lowlink[v] = preorder[v]
                for w in adj[v]:
                    if w not in found:
                        if preorder[w] > preorder[v]:
                            lowlink[v] = min(lowlink[v], lowlink[w])
                        else:
                            lowlink[v] = min(lowlink[v], preorder[w])
                queue.pop()
                if lowlink[v] == preorder[v]:
                    scc = [v]
                    while scc_queue and preorder[scc_queue[-1]] > preorder[v]:
                        scc.append(scc_queue.pop())
                    found.update(scc)
                    out.append(scc)
                else:
                    scc_queue.append(v)
    return out


def scc_labels_reference(src, dst, act, n_nodes: int):
    """compute_SCC_and_Clusters (utils.py:30-52) on the active edges.  Returns (labels i64[N], n_comp)."""
    src, dst = np.asarray(src), np.asarray(dst)
    a = np.flatnonzero(np.asarray(act) != 0)
    sccs = _tarjan_networkx_order(src[a], dst[a])
    sccs = sorted(sccs, key=len)                     # stable, ascending size (utils.py:31)
    labels = np.full(n_nodes, -1, dtype=np.int64)
    k = 0
    for s in sccs:
        labels[np.asarray(s, dtype=np.int64)] = k
        k += 1
    for i in range(n_nodes):                          # isolated nodes appended in index order (utils.py:34-42)
        if labels[i] < 0:
            labels[i] = k
            k += 1
    return labels, k


def scc_partition_canonical(src, dst, act, n_nodes: int):
    """Same partition, canonical numbering: label = smallest node id of the component."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    src, dst = np.asarray(src), np.asarray(dst)
    a = np.flatnonzero(np.asarray(act) != 0)
    g = coo_matrix((np.ones(a.size, dtype=np.int8), (src[a], dst[a])), shape=(n_nodes, n_nodes)).tocsr()
    _, lab = connected_components(g, directed=True, connection="strong")
    first = np.full(lab.max() + 1 if n_nodes else 0, n_nodes, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(n_nodes))
    return first[lab]


def reverse_edge_map(src, dst, n_nodes: int):
    """rev[e] = index of edge (dst[e], src[e]) or -1.  Edges are assumed unique."""
    src, dst = np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64)
    key = src * n_nodes + dst
    order = np.argsort(key, kind="stable")
    skey = key[order]
    want = dst * n_nodes + src
    pos = np.searchsorted(skey, want)
    pos = np.minimum(pos, max(skey.size - 1, 0))
    ok = skey[pos] == want if skey.size else np.zeros(0, dtype=bool)
    return np.where(ok, order[pos], -1)


# --------------------------------------------------------------------------------------
# sequential mirrors (small graphs)
# --------------------------------------------------------------------------------------
def cut_sequential(src, dst, act):
    """remove_edges_single_direction (utils.py:125-142)."""
    act = np.array(act, dtype=np.int64, copy=True)
    active = {(int(src[e]), int(dst[e])) for e in np.flatnonzero(act)}
    for e in np.flatnonzero(act):
        if (int(dst[e]), int(src[e])) not in active:
            act[e] = 0
    return act


def prune_sequential(src, dst, act, prob, num_cameras: int, n_nodes: int):
    """pruning (utils.py:161-188,277-317).  Returns None where the reference returns [] (no violation)."""
    src, dst = np.asarray(src), np.asarray(dst)
    prob = np.asarray(prob)
    act = np.array(act, dtype=np.int64, copy=True)
    fo = np.bincount(src, weights=act, minlength=n_nodes)
    fi = np.bincount(dst, weights=act, minlength=n_nodes)
    vo, vi = np.flatnonzero(fo > num_cameras - 1), np.flatnonzero(fi > num_cameras - 1)
    if vo.size == 0 and vi.size == 0:
        return None
    while True:
        rem = []
        for n in vo:                                   # picks all come from the same snapshot of act
            pos = np.flatnonzero((src == n) & (act == 1))
            rem.append(pos[np.argmin(prob[pos])])      # first index on ties (torch.argmin)
        for n in vi:
            pos = np.flatnonzero((dst == n) & (act == 1))
            rem.append(pos[np.argmin(prob[pos])])
        act[np.asarray(rem, dtype=np.int64)] = 0
        fo = np.bincount(src, weights=act, minlength=n_nodes)
        fi = np.bincount(dst, weights=act, minlength=n_nodes)
        vo, vi = np.flatnonzero(fo > num_cameras - 1), np.flatnonzero(fi > num_cameras - 1)
        if vo.size == 0 and vi.size == 0:
            return act


def split_sequential(src, dst, act, prob, num_cameras: int, n_nodes: int, labels=None):
    """splitting (utils.py:54-123): one cluster, one probability value per SCC recomputation."""
    src, dst = np.asarray(src), np.asarray(dst)
    prob = np.asarray(prob)
    act = np.array(act, dtype=np.int64, copy=True)
    if labels is None:
        labels, _ = scc_labels_reference(src, dst, act, n_nodes)
    while True:
        big = np.flatnonzero(np.bincount(labels) > num_cameras)
        if big.size == 0:
            return act
        l = big[0]
        while True:
            members = labels == l
            a = np.flatnonzero(act == 1)
            touch = a[members[src[a]] | members[dst[a]]]          # either endpoint in the cluster (utils.py:71)
            m = prob[touch].min()
            act[prob == m] = 0                                     # global float equality (utils.py:96-98)
            labels, _ = scc_labels_reference(src, dst, act, n_nodes)
            if not (np.bincount(labels)[l] > num_cameras):         # l re-read in the NEW numbering (utils.py:112)
                break


def post_processing_sequential(src, dst, pred, prob, num_cameras, n_nodes, cutting=True, pruning=True, splitting=True):
    """inference.post_processing (inference.py:70-169).  Returns (labels i64[N], predictions i64[E])."""
    act = np.array(pred, dtype=np.int64, copy=True)
    if cutting:
        act = cut_sequential(src, dst, act)
    if pruning:
        r = prune_sequential(src, dst, act, prob, num_cameras, n_nodes)
        if r is not None:
            act = r
    if cutting:
        act = cut_sequential(src, dst, act)
    if splitting:
        act = split_sequential(src, dst, act, prob, num_cameras, n_nodes)
    labels, _ = scc_labels_reference(src, dst, act, n_nodes)
    return labels, act


# --------------------------------------------------------------------------------------
# parallel-round restatement (large graphs)
# --------------------------------------------------------------------------------------
def cut_rounds(act, rev):
    act = np.asarray(act).astype(bool)
    return act & np.where(rev >= 0, act[np.maximum(rev, 0)], False)


def _segment_argmin(keys_node, prob, eids, n_nodes):
    """For each node, the edge id with min (prob, eid) among the given (node, prob, eid) triples; -1 if none."""
    best = np.full(n_nodes, -1, dtype=np.int64)
    if eids.size == 0:
        return best
    order = np.lexsort((eids, prob, keys_node))
    kn = keys_node[order]
    first = np.ones(kn.size, dtype=bool)
    first[1:] = kn[1:] != kn[:-1]
    best[kn[first]] = eids[order][first]
    return best


def prune_rounds(src, dst, act, prob, num_cameras, n_nodes):
    """Returns (act, changed)."""
    src, dst = np.asarray(src), np.asarray(dst)
    act = np.array(act, dtype=bool, copy=True)
    changed = False
    while True:
        a = np.flatnonzero(act)
        fo = np.bincount(src[a], minlength=n_nodes)
        fi = np.bincount(dst[a], minlength=n_nodes)
        vo, vi = fo > num_cameras - 1, fi > num_cameras - 1
        if not vo.any() and not vi.any():
            return act, changed
        changed = True
        ao = a[vo[src[a]]]
        ai = a[vi[dst[a]]]
        bo = _segment_argmin(src[ao], prob[ao], ao, n_nodes)
        bi = _segment_argmin(dst[ai], prob[ai], ai, n_nodes)
        rem = np.concatenate([bo[bo >= 0], bi[bi >= 0]])
        act[rem] = False


def split_rounds(src, dst, act, prob, num_cameras, n_nodes, report_ties=False):
    """SPLITTING as rounds: every oversized cluster drops its minimum-probability edge(s) in the same round.

    Equal to the reference's one-cluster-at-a-time loop (``split_sequential``) whenever no round has a *cross-cluster tie*: a
    minimum m_A of an oversized cluster A that is also the probability of an active edge touching ANOTHER oversized cluster B
    with m_A > m_B.  The reference compares floats with ``==`` globally (utils.py:96-98), so in its order B may lose that edge
    while A is processed and no longer need to drop its own minimum, whereas here B drops its minimum in the same round.
    (Equal minima of two clusters are harmless, and so are ties with edges of clusters that are not oversized.)  Exact fp32
    ties between different edges' softmax outputs do not occur in generic inputs; ``report_ties=True`` also returns the number
    of rounds in which the condition above held, i.e. in which the equivalence is not guaranteed."""
    src, dst = np.asarray(src), np.asarray(dst)
    prob = np.asarray(prob)
    act = np.array(act, dtype=bool, copy=True)
    tie_rounds = 0
    while True:
        lab = scc_partition_canonical(src, dst, act, n_nodes)
        size = np.bincount(lab, minlength=n_nodes)
        big = size > num_cameras
        if not big.any():
            return (act, tie_rounds) if report_ties else act
        a = np.flatnonzero(act)
        m = np.full(n_nodes, np.inf, dtype=np.float64)
        ls, ld = lab[src[a]], lab[dst[a]]
        s_ok, d_ok = big[ls], big[ld]
        np.minimum.at(m, ls[s_ok], prob[a][s_ok])
        np.minimum.at(m, ld[d_ok], prob[a][d_ok])
        vals = np.unique(m[np.isfinite(m)]).astype(prob.dtype)
        if report_ties and vals.size > 1:
            pa = prob[a]
            hit = np.isin(pa, vals)                    # active edges whose probability is some cluster's minimum ...
            tie = (hit & s_ok & (pa > m[ls])) | (hit & d_ok & (pa > m[ld]))   # ... but not the minimum of an oversized cluster they touch
            tie_rounds += bool(tie.any())
        act &= ~np.isin(prob, vals)                    # every edge anywhere with prob in {m_l}


def post_processing_rounds(src, dst, pred, prob, num_cameras, n_nodes, cutting=True, pruning=True, splitting=True,
                           numbering="canonical"):
    src, dst = np.asarray(src), np.asarray(dst)
    act = np.asarray(pred) != 0
    rev = reverse_edge_map(src, dst, n_nodes) if cutting else None
    if cutting:
        act = cut_rounds(act, rev)
    if pruning:
        act, _ = prune_rounds(src, dst, act, prob, num_cameras, n_nodes)
    if cutting:
        act = cut_rounds(act, rev)
    if splitting:
        act = split_rounds(src, dst, act, prob, num_cameras, n_nodes)
    if numbering == "reference":
        labels, _ = scc_labels_reference(src, dst, act, n_nodes)
    else:
        labels = scc_partition_canonical(src, dst, act, n_nodes)
    return labels, act.astype(np.int64)


# --------------------------------------------------------------------------------------
# synthetic predicted graphs
# --------------------------------------------------------------------------------------
def planted_prediction_graph(n_nodes: int, num_cameras: int, seed: int, n_extra_per_node: float = 2.0,
                             flip_on: float = 0.02, flip_off: float = 0.02, single_dir: float = 0.01,
                             dense: bool = False):
    """Planted clusters (<=1 node per camera) + noise edges.  Returns src,dst i64[E], prob1 f32[E], pred i64[E], cam.

    ``dense=True`` builds the full cross-camera edge set (small N); otherwise a sparse predicted graph
    with symmetric structure: all intra-cluster pairs plus ``n_extra_per_node`` random inter-cluster pairs.
    """
    rng = np.random.default_rng(seed)
    cam = (np.arange(n_nodes) * num_cameras // n_nodes).astype(np.int64)
    # identities: greedily draw one node per camera subset
    ident = np.full(n_nodes, -1, dtype=np.int64)
    per_cam = [list(rng.permutation(np.flatnonzero(cam == c))) for c in range(num_cameras)]
    k = 0
    while any(per_cam):
        avail = [c for c in range(num_cameras) if per_cam[c]]
        sz = rng.integers(1, len(avail) + 1)
        for c in rng.choice(avail, size=sz, replace=False):
            ident[per_cam[c].pop()] = k
        k += 1
    if dense:
        s, d = np.nonzero(cam[:, None] != cam[None, :])
    else:
        order = np.argsort(ident, kind="stable")
        grp = ident[order]
        starts = np.flatnonzero(np.r_[True, grp[1:] != grp[:-1]])
        sizes = np.diff(np.r_[starts, grp.size])
        ps, pd = [], []
        for k in range(2, int(sizes.max()) + 1 if sizes.size else 2):       # vectorised per cluster size (<= num_cameras)
            st = starts[sizes == k]
            if st.size == 0:
                continue
            mem = order[st[:, None] + np.arange(k)[None, :]]                # [n_clusters_of_size_k, k]
            i, j = np.nonzero(~np.eye(k, dtype=bool))
            ps.append(mem[:, i].ravel()); pd.append(mem[:, j].ravel())
        n_extra = int(n_nodes * n_extra_per_node / 2)
        u = rng.integers(0, n_nodes, n_extra); v = rng.integers(0, n_nodes, n_extra)
        ok = cam[u] != cam[v]
        u, v = u[ok], v[ok]
        ps += [u, v]; pd += [v, u]
        s, d = np.concatenate(ps), np.concatenate(pd)
        key = np.unique(s * n_nodes + d)               # unique + lexicographic (row-major) order
        s, d = key // n_nodes, key % n_nodes
    same = ident[s] == ident[d]
    E = s.size
    prob = np.where(same, rng.uniform(0.55, 1.0, E), rng.uniform(0.0, 0.45, E))
    flip = rng.random(E)
    prob = np.where(~same & (flip < flip_on), rng.uniform(0.5, 0.6, E), prob)
    prob = np.where(same & (flip < flip_off), rng.uniform(0.3, 0.5, E), prob)
    sd_ = rng.random(E) < single_dir
    prob = np.where(sd_ & ~same, rng.uniform(0.5, 0.7, E), prob)
    prob = prob.astype(np.float32)
    pred = (prob > 0.5).astype(np.int64)
    return s.astype(np.int64), d.astype(np.int64), prob, pred, cam
