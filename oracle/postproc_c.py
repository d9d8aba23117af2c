"""ctypes loader of the plain-C post-processing oracle (oracle/postproc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Same functions, argument meaning and results as the ``*_rounds`` / ``scc_labels_reference`` functions of
``oracle/postproc_oracle.py``; used where the Python restatement is too slow (10^5 .. 10^8 edges).  The shared object is built
by ``__graft_entry__.build()`` (``build_library()`` below) into ``oracle/_build/`` and is never loaded by the product.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(_HERE, "postproc_oracle.c")
LIB_PATH = os.path.join(_HERE, "_build", "libpostproc_oracle.so")
_lib = None


def build_library(force: bool = False) -> str:
    """gcc -O2 -shared -fPIC oracle/postproc_oracle.c -> oracle/_build/libpostproc_oracle.so (rebuilt when the source is newer)."""
    if force or not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SOURCE):
        os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
        tmp = LIB_PATH + ".tmp%d" % os.getpid()
        try:
            subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Wextra", "-shared", "-fPIC", SOURCE, "-o", tmp, "-lm"], check=True)
            os.replace(tmp, LIB_PATH)                  # atomic: parallel test workers may race here
        except (OSError, subprocess.CalledProcessError):
            if force or not os.path.isfile(LIB_PATH):  # no compiler and nothing prebuilt: the caller cannot check anything
                raise
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build_library())
        p, i64, i32 = C.c_void_p, C.c_int64, C.c_int
        l.po_scc_labels.argtypes = [p, p, p, i64, i64, i32, p, C.POINTER(i64)]
        l.po_reverse_map.argtypes = [p, p, i64, i64, p]
        l.po_cut.argtypes = [p, p, p, i64, i64]
        l.po_prune.argtypes = [p, p, p, p, i64, i64, i32, C.POINTER(i32)]
        l.po_split.argtypes = [p, p, p, p, i64, i64, i32]
        l.po_split_sequential.argtypes = [p, p, p, p, i64, i64, i32]
        l.po_split_hybrid.argtypes = [p, p, p, p, i64, i64, i32, p]
        l.po_post_processing.argtypes = [p, p, p, p, i64, i64, i32, i32, i32, i32, i32, p, p]
        for f in (l.po_scc_labels, l.po_reverse_map, l.po_cut, l.po_prune, l.po_split, l.po_split_sequential, l.po_split_hybrid, l.po_post_processing):
            f.restype = i32
        _lib = l
    return _lib


def _i64(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int64)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _check(rc, what):
    if rc != 0:
        raise RuntimeError("%s failed in the C oracle with status %d (1 = allocation, 2 = bad argument)" % (what, rc))


def scc_labels(src, dst, act, n_nodes: int, numbering: str = "reference"):
    """compute_SCC_and_Clusters (utils.py:30-52) on the active edges -> (labels i64[N], n_components)."""
    src, dst, act = _i64(src), _i64(dst), _i64(act)
    labels = np.empty(n_nodes, dtype=np.int64)
    n = C.c_int64(0)
    _check(lib().po_scc_labels(src.ctypes.data, dst.ctypes.data, act.ctypes.data, src.size, n_nodes,
                               1 if numbering == "reference" else 0, labels.ctypes.data, C.byref(n)), "po_scc_labels")
    return labels, int(n.value)


def reverse_edge_map(src, dst, n_nodes: int):
    src, dst = _i64(src), _i64(dst)
    rev = np.empty(src.size, dtype=np.int64)
    _check(lib().po_reverse_map(src.ctypes.data, dst.ctypes.data, src.size, n_nodes, rev.ctypes.data), "po_reverse_map")
    return rev


def cut(src, dst, act, n_nodes: int):
    src, dst, act = _i64(src), _i64(dst), _i64(act).copy()
    _check(lib().po_cut(src.ctypes.data, dst.ctypes.data, act.ctypes.data, src.size, n_nodes), "po_cut")
    return act


def prune(src, dst, act, prob, num_cameras: int, n_nodes: int):
    """-> (act i64[E], changed): ``changed`` False is the case in which the reference returns []."""
    src, dst, act, prob = _i64(src), _i64(dst), _i64(act).copy(), _f32(prob)
    changed = C.c_int(0)
    _check(lib().po_prune(src.ctypes.data, dst.ctypes.data, act.ctypes.data, prob.ctypes.data, src.size, n_nodes, num_cameras,
                          C.byref(changed)), "po_prune")
    return act, bool(changed.value)


def split(src, dst, act, prob, num_cameras: int, n_nodes: int):
    src, dst, act, prob = _i64(src), _i64(dst), _i64(act).copy(), _f32(prob)
    _check(lib().po_split(src.ctypes.data, dst.ctypes.data, act.ctypes.data, prob.ctypes.data, src.size, n_nodes, num_cameras),
           "po_split")
    return act


def split_sequential(src, dst, act, prob, num_cameras: int, n_nodes: int):
    """splitting (utils.py:54-123) in the reference's own order, one cluster and one SCC pass per dropped value: the exact
    semantics under probability ties, for graphs up to ~10^5 active edges."""
    src, dst, act, prob = _i64(src), _i64(dst), _i64(act).copy(), _f32(prob)
    _check(lib().po_split_sequential(src.ctypes.data, dst.ctypes.data, act.ctypes.data, prob.ctypes.data, src.size, n_nodes,
                                     num_cameras), "po_split_sequential")
    return act


def split_hybrid(src, dst, act, prob, num_cameras: int, n_nodes: int):
    """SPLITTING in the reference's order where probability ties make the order matter, all clusters per iteration elsewhere
    (design study, see the C source).  -> (act, {'tie_values', 'iterations', 'tainted_steps'})."""
    src, dst, act, prob = _i64(src), _i64(dst), _i64(act).copy(), _f32(prob)
    stats = np.zeros(3, dtype=np.int64)
    _check(lib().po_split_hybrid(src.ctypes.data, dst.ctypes.data, act.ctypes.data, prob.ctypes.data, src.size, n_nodes,
                                 num_cameras, stats.ctypes.data), "po_split_hybrid")
    return act, dict(tie_values=int(stats[0]), iterations=int(stats[1]), tainted_steps=int(stats[2]))


def post_processing(src, dst, pred, prob, num_cameras: int, n_nodes: int, cutting=True, pruning=True, splitting=True,
                    numbering: str = "canonical"):
    """inference.post_processing (inference.py:70-169) -> (labels i64[N], predictions i64[E])."""
    src, dst, pred, prob = _i64(src), _i64(dst), _i64(pred), _f32(prob)
    labels = np.empty(n_nodes, dtype=np.int64)
    act = np.empty(src.size, dtype=np.int64)
    _check(lib().po_post_processing(src.ctypes.data, dst.ctypes.data, pred.ctypes.data, prob.ctypes.data, src.size, n_nodes,
                                    num_cameras, int(bool(cutting)), int(bool(pruning)), int(bool(splitting)),
                                    1 if numbering == "reference" else 0, labels.ctypes.data, act.ctypes.data),
           "po_post_processing")
    return labels, act
