"""Evaluation counts of inference.py on the GPU.

``compute_P_R_F`` mirrors inference.py:20-66 (same name, arguments and return tuple).  The clustering scores the reference
takes from ``sklearn.metrics`` on ``(ID_GT, ID_pred)`` (inference.py:507-519) are functions of the contingency table of the two
labelings; the table (integer counts, exact) is built on the device, the scores are evaluated from it on the host in fp64 with
the formulas of scikit-learn 0.24.2 (env_gnn.yml:107): same function names, so ``from gcn_mtmc_b200 import evaluation as
metrics`` binds them unchanged.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .graph import current_stream_ptr, workspace

_KIND = {torch.uint8: 0, torch.bool: 0, torch.int64: 1, torch.float32: 2}


def _flat_for_kernel(t):
    t = t.reshape(-1)
    if t.dtype not in _KIND:
        t = t.to(torch.float32 if t.is_floating_point() else torch.int64)
    return t.contiguous()


def edge_confusion(preds: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """int64 [3,3] on the device: rows = label class, columns = prediction class (0, 1, anything else)."""
    if not (preds.is_cuda and labels.is_cuda):
        raise RuntimeError("edge_confusion needs CUDA tensors: the B200 path has no CPU fallback")
    p, l = _flat_for_kernel(preds), _flat_for_kernel(labels)
    if p.numel() != l.numel():
        raise ValueError("preds and labels must have the same number of elements")
    dev = p.device
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
    counts = torch.empty(9, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().mpn_edge_confusion(p.data_ptr(), _KIND[p.dtype], l.data_ptr(), _KIND[l.dtype], p.numel(),
                                                 counts.data_ptr(), current_stream_ptr(dev)))
    return counts.view(3, 3)


def compute_P_R_F(preds, labels):
    """inference.py:20-66: (TP, FP, TN, FN, P, R, F, precision_class0, precision_class1), tensors on the device of ``preds``."""
    dev = preds.device
    c = edge_confusion(preds, labels).cpu()
    n0, n1 = int(c[0].sum()), int(c[1].sum())                    # len(index_label_0), len(index_label_1)
    tn, fp, fn, tp = int(c[0, 0]), int(c[0, 1]), int(c[1, 0]), int(c[1, 1])
    i64 = dict(dtype=torch.int64, device=dev)
    TP, FP, TN, FN = (torch.tensor(v, **i64) for v in (tp, fp, tn, fn))
    zero = torch.tensor(0.0, device=dev)
    # the two "precisions" are per-class recalls in percent (inference.py:26-37)
    precision_class1 = [zero if tp == 0 else (TP / n1) * 100.0]
    precision_class0 = [zero if tn == 0 else (TN / n0) * 100.0]
    P = TP / (TP + FP) if (tp + fp) != 0 else zero
    R = TP / (TP + FN) if (tp + fn) != 0 else zero
    F = 2 * (P * R) / (P + R) if float(P + R) != 0 else zero
    return TP, FP, TN, FN, P, R, F, precision_class0, precision_class1


# ---------------------------------------------------------------------------------------------- clustering scores
class Contingency:
    """Contingency table of two labelings (sklearn.metrics.cluster.contingency_matrix, classes in sorted order)."""

    def __init__(self, labels_true, labels_pred, device=None):
        a = torch.as_tensor(np.asarray(labels_true) if not torch.is_tensor(labels_true) else labels_true).reshape(-1)
        b = torch.as_tensor(np.asarray(labels_pred) if not torch.is_tensor(labels_pred) else labels_pred).reshape(-1)
        if a.numel() != b.numel():
            raise ValueError("labels_true and labels_pred must have the same length")
        if device is None:
            device = a.device if a.is_cuda else (b.device if b.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("the contingency table is built on a CUDA device: the B200 path has no CPU fallback")
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        self.n = int(a.numel())
        if self.n == 0:
            self.rows = self.cols = self.vals = np.zeros(0, dtype=np.int64)
            self.a = self.b = np.zeros(0, dtype=np.int64)
            return
        ca, ia = torch.unique(a.to(dev), return_inverse=True)            # sorted classes -> compact ids
        cb, ib = torch.unique(b.to(dev), return_inverse=True)
        Ka, Kb = int(ca.numel()), int(cb.numel())
        ia, ib = ia.to(torch.int64).contiguous(), ib.to(torch.int64).contiguous()
        lib = _lib.lib()
        ws = workspace("contingency", dev, lib.mpn_contingency_workspace_bytes(self.n))
        rows, cols, vals = (torch.empty(self.n, dtype=torch.int64, device=dev) for _ in range(3))
        rs, cs = torch.empty(Ka, dtype=torch.int64, device=dev), torch.empty(Kb, dtype=torch.int64, device=dev)
        nnz = C.c_int64(0)
        with torch.cuda.device(dev):
            _lib.check(lib.mpn_contingency(ia.data_ptr(), ib.data_ptr(), self.n, Ka, Kb, rows.data_ptr(), cols.data_ptr(),
                                           vals.data_ptr(), C.byref(nnz), rs.data_ptr(), cs.data_ptr(), ws.data_ptr(), ws.numel(),
                                           current_stream_ptr(dev)))
        k = int(nnz.value)
        r, c, v = rows[:k].cpu().numpy(), cols[:k].cpu().numpy(), vals[:k].cpu().numpy()
        order = np.lexsort((c, r))                                       # row-major, the order scipy's find() yields
        self.rows, self.cols, self.vals = r[order], c[order], v[order]
        self.a, self.b = rs.cpu().numpy(), cs.cpu().numpy()              # marginals (class counts)

    # -- pieces, as sklearn/metrics/cluster/_supervised.py (0.24.2) computes them
    @staticmethod
    def _entropy(counts):
        pi = counts.astype(np.float64)
        pi = pi[pi > 0]
        if pi.size <= 1:
            return 0.0 if pi.size == 1 else 1.0
        s = np.sum(pi)
        return float(-np.sum((pi / s) * (np.log(pi) - math.log(s))))

    def mutual_info(self):
        if self.n == 0:
            return 0.0
        nz = self.vals.astype(np.float64)
        total = float(self.vals.sum())
        pi, pj = self.a, self.b
        log_nm = np.log(nz)
        nm = nz / total
        outer = pi.take(self.rows).astype(np.int64, copy=False) * pj.take(self.cols).astype(np.int64, copy=False)
        log_outer = -np.log(outer) + math.log(pi.sum()) + math.log(pj.sum())
        mi = nm * (log_nm - math.log(total)) + nm * log_outer
        mi = np.where(np.abs(mi) < np.finfo(mi.dtype).eps, 0.0, mi)
        return float(np.clip(mi.sum(), 0.0, None))

    def expected_mutual_info(self):
        a = np.ascontiguousarray(self.a, dtype=np.int64)
        b = np.ascontiguousarray(self.b, dtype=np.int64)
        return float(_lib.lib().mpn_expected_mutual_information_host(a.ctypes.data, a.size, b.ctypes.data, b.size, self.n))

    def pair_confusion(self):
        n = self.n
        vals = self.vals.astype(np.int64)                                 # n <= 2^31 tracklets: every product fits int64
        sum_squares = int((vals * vals).sum())
        ck = int((vals * self.b[self.cols]).sum()) if vals.size else 0     # contingency.dot(n_k).sum()
        cc = int((vals * self.a[self.rows]).sum()) if vals.size else 0     # contingency.T.dot(n_c).sum()
        c11 = sum_squares - n
        c01 = ck - sum_squares
        c10 = cc - sum_squares
        c00 = n * n - c01 - c10 - sum_squares
        return c00, c01, c10, c11                                         # tn, fp, fn, tp

    # -- scores
    def adjusted_rand_score(self):
        tn, fp, fn, tp = self.pair_confusion()
        if fn == 0 and fp == 0:
            return 1.0
        return 2.0 * (tp * tn - fn * fp) / ((tp + fn) * (fn + tn) + (tp + fp) * (fp + tn))

    def adjusted_mutual_info_score(self):
        ka, kb = self.a.size, self.b.size
        if (ka == kb == 1) or (ka == kb == 0):
            return 1.0
        mi, emi = self.mutual_info(), self.expected_mutual_info()
        h_true, h_pred = self._entropy(self.a), self._entropy(self.b)
        denominator = 0.5 * (h_true + h_pred) - emi                     # average_method='arithmetic'
        eps = np.finfo("float64").eps
        denominator = min(denominator, -eps) if denominator < 0 else max(denominator, eps)
        return float((mi - emi) / denominator)

    def homogeneity_completeness_v_measure(self, beta=1.0):
        if self.n == 0:
            return 1.0, 1.0, 1.0
        entropy_c, entropy_k = self._entropy(self.a), self._entropy(self.b)
        mi = self.mutual_info()
        homogeneity = mi / entropy_c if entropy_c else 1.0
        completeness = mi / entropy_k if entropy_k else 1.0
        if homogeneity + completeness == 0.0:
            v = 0.0
        else:
            v = (1 + beta) * homogeneity * completeness / (beta * homogeneity + completeness)
        return float(homogeneity), float(completeness), float(v)


def adjusted_rand_score(labels_true, labels_pred):
    return Contingency(labels_true, labels_pred).adjusted_rand_score()


def adjusted_mutual_info_score(labels_true, labels_pred):
    return Contingency(labels_true, labels_pred).adjusted_mutual_info_score()


def homogeneity_score(labels_true, labels_pred):
    return Contingency(labels_true, labels_pred).homogeneity_completeness_v_measure()[0]


def completeness_score(labels_true, labels_pred):
    return Contingency(labels_true, labels_pred).homogeneity_completeness_v_measure()[1]


def v_measure_score(labels_true, labels_pred, beta=1.0):
    return Contingency(labels_true, labels_pred).homogeneity_completeness_v_measure(beta)[2]


def clustering_scores(labels_true, labels_pred):
    """All five scores of inference.py:509-519 from ONE contingency table."""
    c = Contingency(labels_true, labels_pred)
    h, cm, v = c.homogeneity_completeness_v_measure()
    return {"adjusted_rand_score": c.adjusted_rand_score(), "adjusted_mutual_info_score": c.adjusted_mutual_info_score(),
            "homogeneity_score": h, "completeness_score": cm, "v_measure_score": v}
