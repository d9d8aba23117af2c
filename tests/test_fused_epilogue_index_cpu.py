"""CPU check of the index arithmetic behind the fused distance epilogue of the Gram GEMM (csrc/gram_ef.cu).

The kernel never looks an edge id up: for graphs whose rows are "all columns but one contiguous gap" it computes
e = rowptr[i] + j - (j past the gap ? gap length : 0), walks only the tiles on or above the diagonal when the row block is the
whole graph, writes the mirrored entries from the same accumulator and does not list tiles without edges.  This test replays that
walk in numpy — the gap table and its two end-point checks (the gap search of ge_center_split_kernel; a binary search here, a
32-ary one there: same answer), the tile list (ge_tile_list_kernel), direct and mirrored entries (the epilogue of
gram_ef_kernel<false>) — on the reference's edge order (inference.py:407-413) and checks that every edge is written exactly once
with the right (row, col) pair.  The shipped kernels themselves are checked against the reference's values on the GPU
(tests/test_gpu_parity.py); the pair-ownership rule of the sharded variant is checked through the library itself
(tests/test_host_cpu.py::test_shared_gram_every_pair_computed_exactly_once)."""
import numpy as np
import pytest

from oracle import mpn_oracle as mo

BM = BN = 128


def gap_table(rowptr, col, n_cols, with_flag=False):
    """The gap search of ge_center_split_kernel: first k with col[beg+k] != k (binary search on the non-decreasing col[beg+k] - k), gap length, and
    the two end-point checks that prove the row is 'all columns but that gap' (else the not_one_gap flag is raised)."""
    gap = np.zeros((rowptr.size - 1, 2), dtype=np.int64)
    not_one_gap = 0
    for r in range(rowptr.size - 1):
        beg, deg = rowptr[r], rowptr[r + 1] - rowptr[r]
        lo, hi = 0, deg
        while lo < hi:
            mid = (lo + hi) >> 1
            if col[beg + mid] == mid:
                lo = mid + 1
            else:
                hi = mid
        gl = n_cols - deg
        gap[r] = (lo, gl)
        ok = gl >= 0 and (lo == deg or (col[beg + lo] == lo + gl and col[beg + deg - 1] == n_cols - 1))
        not_one_gap |= int(not ok)
    return (gap, not_one_gap) if with_flag else gap


def replay(rowptr, gap, row0, M, N, sym):
    """(edge id -> (row, col), times written) produced by the tile walk of the kernel for the A block [row0, row0+M)."""
    writes = {}
    rp = rowptr                                          # local rows
    tiles_run = tiles_skipped = 0
    for by in range((M + BM - 1) // BM):
        for bx in range((N + BN - 1) // BN):
            m0, n0 = by * BM, bx * BN
            if sym and n0 + BN <= m0:
                continue
            rows = np.arange(m0, min(m0 + BM, M))
            no_edges = np.all((gap[rows, 0] <= n0) & (gap[rows, 0] + gap[rows, 1] >= min(n0 + BN, N)))
            if sym and no_edges and n0 >= m0 + BM:
                cols = np.arange(n0, min(n0 + BN, N))
                no_edges = np.all((gap[cols, 0] <= m0) & (gap[cols, 0] + gap[cols, 1] >= min(m0 + BM, M)))
            if no_edges:
                tiles_skipped += 1
                continue
            tiles_run += 1
            for r in rows:
                g0, g1 = gap[r, 0], gap[r, 0] + gap[r, 1]
                for c in range(n0, min(n0 + BN, N)):
                    if c < g0 or c >= g1:
                        e = rp[r] + c - (g1 - g0 if c >= g1 else 0)
                        writes.setdefault(int(e), []).append((row0 + r, c))
                if sym and n0 >= m0 + BM:
                    for c in range(n0, min(n0 + BN, N)):
                        h0, h1 = gap[c, 0], gap[c, 0] + gap[c, 1]
                        if r < h0 or r >= h1:
                            e = rp[c] + r - (h1 - h0 if r >= h1 else 0)
                            writes.setdefault(int(e), []).append((c, r))
    return writes, tiles_run, tiles_skipped


@pytest.mark.parametrize("sizes", [(70, 90, 60, 80), (128, 128, 128), (1, 300, 2), (257,), (130, 140)])
def test_closed_form_edge_ids_cover_every_edge_once(sizes):
    import torch
    cam = torch.from_numpy(np.repeat(np.arange(len(sizes)), sizes))
    ei = mo.cross_camera_edge_index(cam).numpy()
    N, E = int(cam.numel()), ei.shape[1]
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, ei[0] + 1, 1)
    rowptr = np.cumsum(rowptr)
    gap = gap_table(rowptr, ei[1], N)
    starts = np.r_[0, np.cumsum(sizes)]
    for c, (a, b) in enumerate(zip(starts[:-1], starts[1:])):               # the gap is the node's own camera segment
        if E:
            assert np.all(gap[a:b, 1] == b - a)
            assert np.all(gap[a:b, 0] == a) or b == N                       # (a trailing gap may also be reported at deg)
    # whole graph, symmetric walk
    writes, run, skipped = replay(rowptr, gap, 0, N, N, sym=True)
    assert sorted(writes) == list(range(E))
    for e, w in writes.items():
        assert len(w) == 1 and w[0] == (ei[0, e], ei[1, e])
    if len(sizes) > 1 and min(sizes) >= 2 * BM:
        assert skipped > 0
    # row-block shards, plain walk
    for r0, r1 in ((0, N // 3), (N // 3, N)):
        if r1 <= r0:
            continue
        local = rowptr[r0:r1 + 1] - rowptr[r0]
        writes, _, _ = replay(local, gap[r0:r1], r0, r1 - r0, N, sym=False)
        e0 = int(rowptr[r0])
        assert sorted(writes) == list(range(int(rowptr[r1]) - e0))
        for e, w in writes.items():
            assert len(w) == 1 and w[0] == (ei[0, e0 + e], ei[1, e0 + e])


def test_same_camera_tiles_are_skipped():
    import torch
    sizes = (256, 384, 256)
    cam = torch.from_numpy(np.repeat(np.arange(len(sizes)), sizes))
    ei = mo.cross_camera_edge_index(cam).numpy()
    N = int(cam.numel())
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, ei[0] + 1, 1)
    rowptr = np.cumsum(rowptr)
    gap = gap_table(rowptr, ei[1], N)
    writes, run, skipped = replay(rowptr, gap, 0, N, N, sym=True)
    assert len(writes) == ei.shape[1]
    assert skipped == 3 + 6 + 3            # upper-triangular tiles inside the cameras: 2x2, 3x3, 2x2 blocks of 128
    assert run + skipped == 7 * 8 // 2


def test_one_gap_shape_is_recognised_exactly():
    """The device-side check must accept exactly the rows that are 'all columns but one contiguous gap' (strictly ascending
    columns are K0's precondition): brute force over every column subset of a small id space."""
    n = 7
    for mask in range(1 << n):
        cols = np.array([c for c in range(n) if mask >> c & 1], dtype=np.int64)
        rowptr = np.array([0, cols.size], dtype=np.int64)
        (gap,), flag = gap_table(rowptr, cols, n, with_flag=True)
        missing = [c for c in range(n) if not mask >> c & 1]
        one_gap = len(missing) == 0 or missing == list(range(missing[0], missing[0] + len(missing)))
        assert (flag == 0) == one_gap, (mask, gap, flag)
        if one_gap:
            expect = np.array([c for c in range(n) if not (gap[0] <= c < gap[0] + gap[1])])
            assert np.array_equal(expect, cols)
