"""Fixture for the exact SPLITTING order at 200 k nodes (tests/test_zz_c_oracle_gpu.py::GPU_CASE).

The checker is the plain-C statement-order oracle (oracle/postproc_oracle.c::po_split_sequential, pinned against the unmodified
reference in tests/test_c_oracle.py); it needs ~8 minutes for this graph (one full SCC pass per dropped value), too long for the
test suites, so its result is stored: the final activity of the edges that were active when SPLITTING started, bit-packed, and
the reference-numbered labels as a SHA-256.  The generator arguments are the test's own (imported from it).

    python tests/golden/make_split200k.py        # writes tests/golden/split200k_sequential.npz
"""
import hashlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import postproc_c as pc
from oracle import postproc_oracle as po

CASE = dict(n_nodes=200000, cams=8, seed=3, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03, single_dir=0.05)


def main():
    c = CASE
    src, dst, prob, pred, _ = po.planted_prediction_graph(c["n_nodes"], c["cams"], c["seed"], n_extra_per_node=c["n_extra_per_node"],
                                                          flip_on=c["flip_on"], flip_off=c["flip_off"], single_dir=c["single_dir"])
    act = pc.cut(src, dst, pred, c["n_nodes"])
    act, _ = pc.prune(src, dst, act, prob, c["cams"], c["n_nodes"])
    act = pc.cut(src, dst, act, c["n_nodes"])
    t = time.time()
    final = pc.split_sequential(src, dst, act, prob, c["cams"], c["n_nodes"])
    print("po_split_sequential: %.0f s" % (time.time() - t))
    labels, n = pc.scc_labels(src, dst, final, c["n_nodes"])
    start_idx = np.flatnonzero(act)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "split200k_sequential.npz")
    np.savez_compressed(out, spec=np.array([c["n_nodes"], c["cams"], c["seed"]], dtype=np.int64),
                        n_edges=np.array([src.size], dtype=np.int64), n_start_active=np.array([start_idx.size], dtype=np.int64),
                        final_bits=np.packbits(final[start_idx].astype(np.uint8)),
                        labels_sha256=np.frombuffer(hashlib.sha256(np.ascontiguousarray(labels, dtype=np.int64).tobytes()).digest(), dtype=np.uint8),
                        n_components=np.array([n], dtype=np.int64))
    print("wrote", out, "active after SPLITTING:", int(final.sum()))


if __name__ == "__main__":
    main()
