"""Parity of the CUDA path (through the C ABI) against the oracle and the reference's golden outputs.  B200 only."""
import copy
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import mpn_oracle as mo
from oracle import postproc_oracle as po
from tests._util import Data, MPN_FILES, POST_FILES, ids, load_mpn_case, same_partition

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def m():
    import gcn_mtmc_b200 as mod
    mod._lib.require_device(0)
    return mod


def dev():
    return torch.device("cuda", 0)


# ---------------------------------------------------------------------------------------------- graph tables
def test_graph_tables_match_numpy(m):
    x, ei, cam, _ = mo.synth_graph(90, 5, 3, D=8)
    g = m.TrackletGraph(ei.to(dev()), 90, chunk=32)
    row, col = ei[0].numpy(), ei[1].numpy()
    rowptr = np.searchsorted(row, np.arange(91))
    assert np.array_equal(g.rowptr.cpu().numpy(), rowptr)
    assert np.array_equal(g.col.cpu().numpy()[:row.size], col)
    deg = np.diff(rowptr)
    nt = (deg + 31) // 32
    assert int(g.n_tasks.item()) == nt.sum()
    assert np.array_equal(g.taskptr.cpu().numpy(), np.r_[0, np.cumsum(nt)])
    assert np.array_equal(g.task_row.cpu().numpy()[:nt.sum()], np.repeat(np.arange(90), nt))
    assert g.perm is None


def test_graph_from_cameras_matches_edge_index_path(m):
    """Row f1: tables built on the device from camera ids == tables built from the reference's int64 edge_index."""
    cam = torch.tensor([0] * 31 + [2] * 17 + [3] * 40 + [7] * 25)            # ragged cameras, non-contiguous ids
    ei = mo.cross_camera_edge_index(cam)
    g1 = m.TrackletGraph(ei.to(dev()), cam.numel(), chunk=64)
    g2 = m.TrackletGraph.from_cameras(cam, dev(), chunk=64, materialize_edge_index=True)
    torch.cuda.synchronize()
    assert g2.n_edges == ei.shape[1]
    assert torch.equal(g2.edge_index.cpu(), ei)
    for name in ("rowptr", "taskptr", "n_tasks"):
        assert torch.equal(getattr(g1, name), getattr(g2, name)), name
    assert torch.equal(g1.col[:g1.n_edges], g2.col[:g2.n_edges])
    nt = int(g1.n_tasks.item())
    assert torch.equal(g1.task_row[:nt], g2.task_row[:nt])
    with pytest.raises(ValueError):
        m.TrackletGraph.from_cameras(torch.tensor([0, 1, 0, 1]), dev())
    # row-block shards built from the camera layout == shards cut from the edge list
    for (n0, n1) in [(0, 40), (40, 95), (95, 113)]:
        lo, hi = m.shard_edges(ei, n0, n1)
        gs = m.TrackletGraph.from_cameras(cam, dev(), chunk=64, row_block=(n0, n1), materialize_edge_index=True)
        gr = m.TrackletGraph(ei[:, lo:hi].to(dev()), cam.numel(), chunk=64, row_offset=n0, n_rows=n1 - n0)
        assert gs.n_edges == hi - lo and torch.equal(gs.edge_index.cpu(), ei[:, lo:hi])
        assert torch.equal(gs.rowptr, gr.rowptr) and torch.equal(gs.col[:gs.n_edges], gr.col[:gr.n_edges])
        assert torch.equal(gs.taskptr, gr.taskptr)
    # forward through data.mpn_graph (no edge_index on the data object)
    params = mo.shipped_model_params(1, 1, 64, (48,))
    sd = mo.init_weights(params, "resnet101", 3)
    x = torch.nn.functional.normalize(torch.randn(cam.numel(), 64, generator=torch.Generator().manual_seed(0)), dim=0)
    ea = mo.edge_features(x, ei)
    ref, _ = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    g3 = m.TrackletGraph.from_cameras(cam, dev())
    ea_dev = m.edge_features(x.to(dev()), None, graph=g3, use_tensor_cores=False)
    assert np.allclose(ea_dev.cpu().numpy(), ea.numpy(), rtol=3e-6, atol=3e-6)
    out, _ = net(Data(x=x.to(dev()), edge_attr=ea_dev, mpn_graph=g3))
    assert (out["classified_edges"][0].cpu().double() - ref[0]).abs().max().item() <= 1e-4 * ref[0].abs().max().item()


def test_graph_unsorted_and_invalid(m):
    x, ei, cam, _ = mo.synth_graph(40, 4, 5, D=8)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))
    g = m.TrackletGraph(ei[:, perm].to(dev()), 40)
    assert g.perm is not None
    assert torch.equal(ei[:, perm][:, g.perm.cpu()], ei)
    with pytest.raises(ValueError):
        m.TrackletGraph(torch.cat([ei, ei[:, :3]], dim=1).to(dev()), 40)          # duplicates
    bad = ei.clone(); bad[1, 7] = 99
    with pytest.raises(m._lib.MpnError):
        m.TrackletGraph(bad.to(dev()), 40)                                         # node id out of range


def test_graph_deferred_validation(m):
    """validate='deferred': no host round trip in the build; same tables; an invalid edge list leaves an EMPTY graph on the
    device (sweeps touch nothing) and raises at .validate() what the synchronous build raises immediately."""
    x, ei, cam, _ = mo.synth_graph(60, 4, 5, D=64)
    g0 = m.TrackletGraph(ei.to(dev()), 60, chunk=32)
    g1 = m.TrackletGraph(ei.to(dev()), 60, chunk=32, validate="deferred")
    assert g1.validate() is g1 and g1.validate() is g1                          # idempotent
    for name in ("rowptr", "taskptr", "n_tasks"):
        assert torch.equal(getattr(g0, name), getattr(g1, name)), name
    assert torch.equal(g0.col[:g0.n_edges], g1.col[:g1.n_edges])
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(0))
    gu = m.TrackletGraph(ei[:, perm].to(dev()), 60, validate="deferred")
    params = mo.shipped_model_params(1, 1, 64, (48,))
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(mo.init_weights(params, "resnet101", 3), strict=True)
    net = net.to(dev()).eval()
    ea = torch.rand(ei.shape[1], 2, device=dev())
    net.use_cuda_graph = False
    net(Data(x=x.to(dev()), edge_attr=ea, mpn_graph=gu))                        # enqueued on the emptied tables: must not fault
    torch.cuda.synchronize()
    assert int(gu.n_tasks.item()) == 0 and int(gu.rowptr.abs().sum().item()) == 0
    with pytest.raises(m._lib.UnsortedEdgeIndex):
        gu.validate()
    bad = ei.clone(); bad[1, 7] = 99
    gb = m.TrackletGraph(bad.to(dev()), 60, validate="deferred")
    with pytest.raises(m._lib.MpnError):
        gb.validate()
    with pytest.raises(ValueError):
        m.TrackletGraph(ei.to(dev()), 60, validate="later")


# ---------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("shape", [(300, 1024, 2048), (257, 130, 96), (64, 32, 128), (1000, 512, 1024), (33, 7, 50)])
def test_gemm_simt_matches_fp64(m, shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    A, B, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    ref = (A.double() @ B.double().t() + bias.double())
    Ad, Bd, bd = A.to(dev()), B.to(dev()), bias.to(dev())
    Cd = torch.empty(M, N, device=dev())
    ws = torch.empty(4096, dtype=torch.uint8, device=dev())
    m._lib.check(m._lib.lib().mpn_gemm_nt(Ad.data_ptr(), Bd.data_ptr(), bd.data_ptr(), Cd.data_ptr(), M, N, K, 0,
                                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    err = (Cd.cpu().double() - ref).abs().max().item()
    assert err <= 5e-7 * K, err                        # fp32 accumulation: ~eps*K typical, max over M*N entries (values ~N(0,1))


@pytest.mark.parametrize("shape", [(300, 1024, 2048), (257, 130, 96), (128, 256, 64), (1000, 512, 1024), (4096, 128, 512),
                                   (77, 64, 2048), (1024, 1024, 2048)])
def test_gemm_tcgen05_3xtf32_matches_fp64(m, shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A, B, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
    ref = (A.double() @ B.double().t() + bias.double())
    Ad, Bd, bd = A.to(dev()), B.to(dev()), bias.to(dev())
    Cd = torch.full((M, N), float("nan"), device=dev())
    L = m._lib.lib()
    ws = torch.empty(L.mpn_gemm_nt_workspace_bytes(M, N, K, 1), dtype=torch.uint8, device=dev())
    m._lib.check(L.mpn_gemm_nt(Ad.data_ptr(), Bd.data_ptr(), bd.data_ptr(), Cd.data_ptr(), M, N, K, 1,
                               ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (Cd.cpu().double() - ref).abs().max().item()
    # 3xTF32 products are fp32-accurate; the tensor core's truncating accumulate adds a bias ~3e-7*K (measured)
    assert err <= 6e-7 * K, err


@pytest.mark.parametrize("N,K,row0,M,scale", [(640, 256, 0, 640, 1.0), (640, 2048, 128, 320, 0.015), (1000, 64, 0, 1000, 3e4),
                                               (300, 96, 77, 100, 1e-6)])
def test_gram_3xfp16_matches_fp64(m, N, K, row0, M, scale):
    """The Gram block of the edge features on fp16 operand planes (x * 2^k = hi + lo, three kind::f16 products): fp32-level
    accuracy at any input magnitude (the planes are scaled from max |X|), symmetric and row-block shapes."""
    g = torch.Generator().manual_seed(N + K + row0)
    X = torch.randn(N, K, generator=g) * scale
    X[5] *= 1e-3                                                              # a row far below the maximum
    Xd = X.to(dev())
    L = m._lib.lib()
    ws = torch.empty(L.mpn_gemm_nt_workspace_bytes(M, N, K, 1), dtype=torch.uint8, device=dev())
    amax = Xd.abs().max().reshape(1)
    ref = X[row0:row0 + M].double() @ X.double().t()
    # The tensor core adds into its fp32 accumulator with truncation: the bias grows with |accumulator| x accumulation steps, so
    # coherent sums (|x|^2, the entries where row == column) carry ~5e-6 relative, incoherent ones stay at the fp32 level.  The
    # edge features never use the coherent entries (node norms come from fp64 row statistics, near-duplicates are recomputed).
    bound = 1e-5 * K * scale * scale
    incoherent = ref.abs() < 0.2 * K * scale * scale
    for am in (amax.data_ptr(), None):
        Cd = torch.full((M, N), float("nan"), device=dev())
        m._lib.check(L.mpn_gram_nt(Xd.data_ptr(), row0, Cd.data_ptr(), M, N, K, am, ws.data_ptr(), ws.numel(),
                                   torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        diff = (Cd.cpu().double() - ref).abs()
        assert diff.max().item() <= bound, (am is None, diff.max().item(), bound)
        assert diff[incoherent].max().item() <= 6e-7 * K * scale * scale, (am is None, diff[incoherent].max().item())


def test_gemm_tcgen05_gram_aliasing(m):
    """A given as a row block of B (the Gram-matrix call of the edge features) shares the split planes."""
    g = torch.Generator().manual_seed(5)
    X = torch.randn(640, 256, generator=g)
    Xd = X.to(dev())
    r0, r1 = 128, 448
    Cd = torch.empty(r1 - r0, 640, device=dev())
    L = m._lib.lib()
    ws = torch.empty(L.mpn_gemm_nt_workspace_bytes(r1 - r0, 640, 256, 1), dtype=torch.uint8, device=dev())
    m._lib.check(L.mpn_gemm_nt(Xd[r0:].data_ptr(), Xd.data_ptr(), None, Cd.data_ptr(), r1 - r0, 640, 256, 1,
                               ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    ref = X[r0:r1].double() @ X.double().t()
    # coherent sums (|x|^2 on the aliased diagonal) see the largest truncation bias: ~1e-6 relative to |a||b| = K
    assert (Cd.cpu().double() - ref).abs().max().item() <= 2e-6 * 256


# ---------------------------------------------------------------------------------------------- edge features
@pytest.mark.parametrize("tc", [False, True], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("path", MPN_FILES, ids=ids(MPN_FILES))
def test_edge_features_match_reference_golden(m, path, tc):
    """Both GEMM paths against the edge features the unmodified reference produced (tests/golden/mpn_*.npz): the fp32 SIMT Gram +
    gather pass, and the tensor-core path (the fused tcgen05 kernel of csrc/gram_ef.cu for these dense cross-camera graphs when
    D % 64 == 0, else the tcgen05 Gram + gather) — the one the benchmark times."""
    g, params, sd, x, ei, _ = load_mpn_case(path)
    out = m.edge_features(x.to(dev()), ei.to(dev()), use_tensor_cores=tc).cpu().numpy()
    assert np.allclose(out, g["edge_attr"], rtol=3e-6, atol=3e-6), np.abs(out - g["edge_attr"]).max()


def test_edge_features_near_duplicates_and_unsorted(m):
    x, ei, cam, ident = mo.synth_graph(64, 4, 21, D=256, planted=True, noise=1e-3)     # near-duplicate embeddings
    ref = mo.edge_features(x, ei, dtype=torch.float64).numpy()
    out = m.edge_features(x.to(dev()), ei.to(dev()), use_tensor_cores=False).cpu().numpy()
    assert np.allclose(out, ref, rtol=2e-5, atol=1e-6)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(1))
    out_p = m.edge_features(x.to(dev()), ei[:, perm].to(dev()), use_tensor_cores=False).cpu().numpy()
    assert np.array_equal(out_p, out[perm.numpy()])


# ---------------------------------------------------------------------------------------------- forward
def run_forward(m, params, sd, x, ei, ea, fuse=False, chunk=None):
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    net.fuse_decisions = fuse
    data = Data(x=x.to(dev()), edge_index=ei.to(dev()), edge_attr=ea.to(dev()))
    if chunk is not None:                                # force the task size (>= 128 selects the tensor-core apply kernel)
        data.mpn_graph = m.TrackletGraph(data.edge_index, x.shape[0], chunk=chunk)
        net.use_cuda_graph = False
    out, h = net(data)
    torch.cuda.synchronize()
    return out["classified_edges"], h, net


@pytest.mark.parametrize("path", MPN_FILES, ids=ids(MPN_FILES))
def test_forward_matches_reference_golden(m, path):
    g, params, sd, x, ei, _ = load_mpn_case(path)
    ea = torch.from_numpy(g["edge_attr"])
    outs, h, net = run_forward(m, params, sd, x, ei, ea, fuse=True)
    assert len(outs) == int(g["n_logits"][0])
    for i, o in enumerate(outs):
        ref = g[f"logits{i}"]
        tol = 1e-4 * np.abs(ref).max()                   # north_star tolerance: 1e-4 relative (to max |logit|)
        assert np.abs(o.cpu().numpy() - ref).max() <= tol, (i, np.abs(o.cpu().numpy() - ref).max(), tol)
    assert np.abs(h.cpu().numpy() - g["h"]).max() <= 1e-4 * max(1.0, np.abs(g["h"]).max())
    last = g[f"logits{len(outs) - 1}"]
    margin = np.abs(last[:, 1] - last[:, 0])
    pred = net.last_pred.cpu().numpy()
    assert not np.any((pred != g["pred"]) & (margin > 1e-4))      # decisions identical outside the 1e-4 band
    assert np.abs(net.last_prob1.cpu().numpy() - g["prob"][:, 1]).max() <= 2e-5
    # the fused probability is ATen's own softmax arithmetic: bit-identical to torch.softmax of the emitted logits, and to mpn_decide
    assert torch.equal(net.last_prob1, torch.softmax(outs[-1], dim=1)[:, 1])
    pred2 = torch.empty_like(net.last_pred)
    prob2 = torch.empty_like(net.last_prob1)
    lg = outs[-1].contiguous()
    m._lib.check(m._lib.lib().mpn_decide(lg.data_ptr(), lg.shape[0], pred2.data_ptr(), prob2.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
    assert torch.equal(prob2, net.last_prob1) and torch.equal(pred2, net.last_pred)


def test_forward_s02_shape_vs_oracle_and_deterministic(m):
    params = mo.shipped_model_params(1, 1)
    x, ei, cam, _ = mo.synth_graph(300, 4, 0, planted=True)
    sd = mo.init_weights(params, "resnet101", 7)
    ea = mo.edge_features(x, ei)
    ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei, ea)
    ref64, _ = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    outs, h, net = run_forward(m, params, sd, x, ei, ea)
    o = outs[0].cpu()
    scale = ref64[0].abs().max().item()
    err = (o.double() - ref64[0]).abs().max().item()
    err_ref = (ref[0].double() - ref64[0]).abs().max().item()
    assert err <= max(1e-4 * scale, err_ref), (err, err_ref)
    assert (h.cpu() - href).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item())
    outs2, h2, _ = run_forward(m, params, sd, x, ei, ea)
    assert torch.equal(outs2[0].cpu(), o) and torch.equal(h2.cpu(), h.cpu())        # bit-reproducible


def test_forward_multi_step_unsorted_edges(m):
    params = mo.shipped_model_params(3, 2, 64, (48, 40))
    x, ei, cam, _ = mo.synth_graph(80, 4, 31, D=64, planted=True)
    sd = mo.init_weights(params, "resnet101", 9)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(2))
    ei_p = ei[:, perm]
    ea_p = mo.edge_features(x, ei_p)
    ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei_p, ea_p, dtype=torch.float64)
    outs, h, _ = run_forward(m, params, sd, x, ei_p, ea_p)
    for o, r in zip(outs, ref):
        assert (o.cpu().double() - r).abs().max().item() <= 1e-4 * r.abs().max().item()
    assert (h.cpu().double() - href).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item())


@pytest.mark.parametrize("L,n_cls,chunk,fuse,agg", [(1, 1, 128, True, "sum"), (3, 2, 128, False, "sum"), (2, 1, 256, True, "sum"),
                                                    (1, 1, 1024, False, "sum"), (2, 1, 128, True, "max"), (3, 1, 256, False, "mean"),
                                                    (2, 2, 32, False, "max"), (2, 1, 64, True, "mean")])
def test_forward_tensor_core_apply_vs_fp64_oracle(m, L, n_cls, chunk, fuse, agg):
    """Tasks of >= 128 edges take the stored-y + tcgen05 apply path (sum|z| + closed-form half in node_finalize): rows of 400
    edges = 3 full batches + a masked tail; a thinned copy adds short rows, empty rows and runs that change row every batch."""
    params = mo.shipped_model_params(L, n_cls, 64, (48, 40))
    params["node_agg_fn"] = agg                                                # models/mpn.py:193-202 (chunk < 128: packed-fp32 kernel)
    x, ei, cam, _ = mo.synth_graph(600, 3, 5, D=64, planted=True)
    sd = mo.init_weights(params, "resnet101", 11, affine_jitter=True)
    keep = torch.rand(ei.shape[1], generator=torch.Generator().manual_seed(3)) < 0.6
    keep &= ~((ei[0] >= 100) & (ei[0] < 140))                                  # rows without edges
    keep &= ~((ei[0] >= 300) & (ei[0] < 330) & (ei[1] % 7 != 0))               # short rows (< 128 edges)
    for ei_t in (ei, ei[:, keep]):
        ea = mo.edge_features(x, ei_t)
        ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei_t, ea, dtype=torch.float64)
        outs, h, net = run_forward(m, params, sd, x, ei_t, ea, fuse=fuse, chunk=chunk)
        for o, r in zip(outs, ref):
            assert (o.cpu().double() - r).abs().max().item() <= 1e-4 * r.abs().max().item()
        assert (h.cpu().double() - href).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item())
        outs2, h2, _ = run_forward(m, params, sd, x, ei_t, ea, fuse=fuse, chunk=chunk)
        assert torch.equal(outs2[-1], outs[-1]) and torch.equal(h2, h)          # bit-reproducible
        if fuse:
            margin = (ref[-1][:, 1] - ref[-1][:, 0]).abs()
            pred_ref = (ref[-1][:, 1] > ref[-1][:, 0]).to(torch.uint8)
            assert not bool(((net.last_pred.cpu() != pred_ref) & (margin > 1e-4)).any())
            prob_ref = torch.softmax(ref[-1], dim=1)[:, 1]
            assert (net.last_prob1.cpu().double() - prob_ref).abs().max().item() <= 5e-6


@pytest.mark.parametrize("re_n,re_e,L,n_cls,chunk,agg", [(True, True, 3, 2, 128, "sum"), (True, False, 2, 1, 128, "sum"),
                                                         (False, True, 3, 1, 256, "max"), (True, True, 1, 1, 128, "sum"),
                                                         (True, True, 4, 4, 32, "mean"), (False, True, 2, 2, None, "sum")])
def test_forward_reattach_initial_features(m, re_n, re_e, L, n_cls, chunk, agg):
    """reattach_initial_nodes / reattach_initial_edges (models/mpn.py:207-215, 283-287): [h0 | h] and [e0 | e] before every
    step, on the packed-fp32 (chunk < 128) and the tensor-core sweeps, against the fp64 oracle."""
    params = mo.shipped_model_params(L, n_cls, 64, (48, 40))
    params["reattach_initial_nodes"], params["reattach_initial_edges"], params["node_agg_fn"] = re_n, re_e, agg
    x, ei, cam, _ = mo.synth_graph(600, 3, 8, D=64, planted=True)
    sd = mo.init_weights(params, "resnet101", 13, affine_jitter=True)
    assert sd["MPNet.edge_model.edge_mlp.fc_layers.0.weight"].shape[1] == (2 if re_n else 1) * 64 + (2 if re_e else 1) * 4
    ea = mo.edge_features(x, ei)
    ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    outs, h, net = run_forward(m, params, sd, x, ei, ea, fuse=True, chunk=chunk)
    assert len(outs) == n_cls
    for o, r in zip(outs, ref):
        assert (o.cpu().double() - r).abs().max().item() <= 1e-4 * r.abs().max().item()
    assert (h.cpu().double() - href).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item())
    margin = (ref[-1][:, 1] - ref[-1][:, 0]).abs()
    assert not bool(((net.last_pred.cpu() != (ref[-1][:, 1] > ref[-1][:, 0]).to(torch.uint8)) & (margin > 1e-4)).any())


def test_forward_computes_edge_features_when_absent(m):
    """data.edge_attr = None: K1 runs inside forward (node encoder on the side stream) — same edge features as the two-call path
    (handed back bit for bit), on the large-graph path, the CUDA-graph path for small graphs and unsorted edges; same logits bit
    for bit where the forward is the same launches (CUDA-graph path, unsorted graphs), within 2e-6 * max|logit| where the first
    encoder BatchNorm's moment sums come from the GEMM epilogue instead of a sweep (dense sorted graphs on the direct path)."""
    params = mo.shipped_model_params(2, 1, 64, (48, 40))
    sd = mo.init_weights(params, "resnet101", 21)
    x, ei, cam, _ = mo.synth_graph(200, 4, 9, D=64, planted=True)
    perm = torch.randperm(ei.shape[1], generator=torch.Generator().manual_seed(4))
    for edges, cuda_graph in ((ei, False), (ei, True), (ei[:, perm], False)):
        net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
        net.load_state_dict(sd, strict=True)
        net = net.to(dev()).eval()
        net.use_cuda_graph = cuda_graph
        xd, eid = x.to(dev()), edges.to(dev())
        ea = m.edge_features(xd, eid)
        o1, h1 = net(Data(x=xd, edge_index=eid, edge_attr=ea))
        d2 = Data(x=xd, edge_index=eid, edge_attr=None)
        o2, h2 = net(d2)
        torch.cuda.synchronize()
        assert torch.equal(d2.edge_attr, ea)
        l1, l2 = o1["classified_edges"][-1], o2["classified_edges"][-1]
        if cuda_graph:
            assert torch.equal(l1, l2) and torch.equal(h1, h2)
        else:
            assert (l1 - l2).abs().max().item() <= 2e-6 * l1.abs().max().item()
            assert (h1 - h2).abs().max().item() <= 2e-6 * max(1.0, h1.abs().max().item())
    d3 = Data(x=x.to(dev()), edge_index=ei.to(dev()))                          # attribute absent altogether
    net(d3)
    assert torch.equal(d3.edge_attr, m.edge_features(x.to(dev()), ei.to(dev())))


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_graph_stream_pipelined_copies_match_direct_forward(m, depth):
    """GraphStream: several graphs of different shapes in flight (H2D / kernels / D2H on three streams) give the same bits as
    one direct forward per graph; buffers are reused every ``depth`` submits."""
    params = mo.shipped_model_params(1, 1, 64, (48, 40))
    sd = mo.init_weights(params, "resnet101", 21)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    net.fuse_decisions = True
    cases = []
    for (N, Cn, seed) in [(200, 4, 1), (2048, 4, 2), (200, 4, 3), (330, 3, 4), (2048, 4, 5), (2048, 4, 6), (200, 4, 7)]:
        x, ei, cam, _ = mo.synth_graph(N, Cn, seed, D=64, planted=True)
        d = Data(x=x.to(dev()), edge_index=ei.to(dev()))
        net(d)
        cases.append((x.pin_memory(), cam.numpy(), net.last_pred.cpu(), net.last_prob1.cpu(), ei.shape[1]))
    gs = m.GraphStream(net, dev(), depth=depth)
    outs = [(torch.zeros(c[4], dtype=torch.uint8).pin_memory(), torch.zeros(c[4], dtype=torch.float32).pin_memory()) for c in cases]
    tickets = [gs.submit(c[0], c[1], o[0], o[1]) for c, o in zip(cases, outs)]
    gs.wait(tickets[0])
    assert torch.equal(outs[0][0], cases[0][2])
    gs.drain()
    for c, o in zip(cases, outs):
        assert torch.equal(o[0], c[2]) and torch.equal(o[1], c[3])
    assert net.fuse_decisions is True
    with pytest.raises(ValueError):
        gs.submit(cases[0][0].clone(), cases[0][1], outs[0][0])                 # not pinned
    with pytest.raises(ValueError):
        gs.submit(cases[0][0], cases[0][1], outs[1][0])                         # wrong number of edges


def test_graph_stream_graph_replay(m):
    """Slot-level CUDA-graph replay: the third and later submits of one signature per slot are replays; same bits as eager."""
    params = mo.shipped_model_params(1, 1, 64, (48, 40))
    sd = mo.init_weights(params, "resnet101", 21)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    net.fuse_decisions = True
    for N in (200, 2048):                                                      # small-graph path and large-graph path
        xs, refs, cam = [], [], None
        for seed in range(6):
            x, ei, cam, _ = mo.synth_graph(N, 4, seed, D=64, planted=True)
            net(Data(x=x.to(dev()), edge_index=ei.to(dev())))
            xs.append(x.pin_memory()); refs.append(net.last_pred.cpu())
        gs = m.GraphStream(net, dev(), depth=2, graph_replay=True)
        outs = [torch.zeros(refs[0].numel(), dtype=torch.uint8).pin_memory() for _ in xs]
        for x, o in zip(xs, outs):                                            # per slot: eager, capture + replay, replay
            gs.submit(x, cam.numpy(), o)
        gs.drain()
        assert all(s.cap is not None for s in gs.slots)
        for o, r in zip(outs, refs):
            assert torch.equal(o, r)


def test_programmatic_dependent_launch(m):
    """Same bits with and without the launch attribute (the default is on): graph tables, edge features, forward (small-graph
    replay and large-graph path)."""
    lib = m._lib.lib()
    assert lib.mpn_set_pdl(-1) == 2, "programmatic dependent launch is the default"
    params = mo.shipped_model_params(2, 1, 64, (48, 40))
    sd = mo.init_weights(params, "resnet101", 23)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    net.fuse_decisions = True
    try:
        for N in (200, 3000):
            x, ei, cam, _ = mo.synth_graph(N, 4, 7, D=64, planted=True)
            outs = []
            for on in (0, 1, 1, 0):
                assert lib.mpn_set_pdl(on) == 1 + on
                g = m.TrackletGraph.from_cameras(cam.numpy(), dev())
                d = Data(x=x.to(dev()), edge_index=None, mpn_graph=g, edge_attr=None)
                net._graphs.clear()                              # the small-graph CUDA graph is re-captured in each mode
                out, h = net(d)
                torch.cuda.synchronize()
                outs.append((d.edge_attr.clone(), out["classified_edges"][-1].clone(), h.clone(), net.last_pred.clone()))
            for o in outs[1:]:
                for a, b in zip(outs[0], o):
                    assert torch.equal(a, b)
    finally:
        lib.mpn_set_pdl(1)


def test_fused_edge_feature_kernel_vs_reference_ops(m):
    """csrc/gram_ef.cu — dense cross-camera graphs: edge features formed in the epilogue of the persistent Gram GEMM — against the
    reference's own operations (F.pairwise_distance / F.cosine_similarity on gathered rows, inference.py:453-456, restated in
    oracle/mpn_oracle.py): unequal cameras, sizes that are not tile multiples, D = 64 .. 2048, near-duplicate rows (refine list),
    a zero row (cosine clamp), row-block shards, graphs given as an int64 edge_index (layout decided on the device) and a graph
    that is NOT dense (falls back to Gram + gather inside the same call).  Gate: rtol = atol = 1e-5; every edge written; run-to-run
    bit-identical."""
    import numpy as np
    for sizes, D, seed in (((70, 90, 60, 80), 64, 1), ((124, 90, 99, 137), 2048, 2), ((300, 260, 200, 141, 123), 128, 3),
                           ((512,) * 8, 256, 4), ((1, 2, 130), 64, 5), ((700, 3), 192, 6)):
        cam = np.repeat(np.arange(len(sizes)), sizes)
        N = int(cam.size)
        gen = torch.Generator().manual_seed(seed)
        x = torch.randn(N, D, generator=gen)
        x[N // 2] = x[3] + 1e-4 * torch.randn(D, generator=gen)            # same-identity pairs: cancellation -> refine list
        x[N - 1] = x[0]
        if N > 200:
            x[7] = 0.0                                                     # |a| = 0: the reference clamps |a||b| at 1e-8
        x = torch.nn.functional.normalize(x, p=2, dim=0)
        x[7 if N > 200 else 0] *= 1.0
        ei = torch.cat([torch.cartesian_prod(torch.nonzero(torch.from_numpy(cam) == c).reshape(-1),
                                             torch.nonzero(torch.from_numpy(cam) != c).reshape(-1))
                        for c in range(len(sizes))], dim=0).t().contiguous()
        ref = mo.edge_features(x, ei)
        xd = x.to(dev())
        for block in (None, (0, N // 3), (N // 3, N)):
            g = m.TrackletGraph.from_cameras(cam, dev(), row_block=block)
            assert g.struct.layout_hint == 1
            got = torch.full((g.n_edges, 2), float("nan"), device=dev())          # every edge must be written
            assert m.edge_features(xd, None, graph=g, out=got) is got
            again = m.edge_features(xd, None, graph=g)
            torch.cuda.synchronize()
            assert torch.isfinite(got).all() and torch.equal(got, again)
            if block is None:
                want = ref
            else:
                sel = (ei[0] >= block[0]) & (ei[0] < block[1])
                want = ref[sel]
            assert np.allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-5), (sizes, D, block, (got.cpu() - want).abs().max())
        # the same graph as an int64 edge_index: the layout is recognised on the device, both launch sets are enqueued
        g2 = m.TrackletGraph(ei.to(dev()), N)
        assert g2.struct.layout_hint == 0
        got2 = m.edge_features(xd, None, graph=g2)
        assert torch.equal(got2, m.edge_features(xd, None, graph=m.TrackletGraph.from_cameras(cam, dev())))
    # a graph that is not "all columns but one gap" takes the Gram + gather pass inside the same call
    x, ei, cam, _ = mo.synth_graph(1100, 5, 4, D=64, planted=True)
    keep = torch.rand(ei.shape[1], generator=torch.Generator().manual_seed(0)) < 0.7
    keep[:5] = torch.tensor([True, False, True, False, True])
    eit = ei[:, keep].contiguous()
    got = m.edge_features(x.to(dev()), eit.to(dev()))
    assert np.allclose(got.cpu().numpy(), mo.edge_features(x, eit).numpy(), rtol=1e-5, atol=1e-5)
    assert torch.equal(got, m.edge_features(x.to(dev()), eit.to(dev()), use_tensor_cores=True))


def test_forward_with_fused_moments_matches_the_sweep(m):
    """MOTMPNet.forward with data.edge_attr = None (K1 inside the forward: the first encoder BatchNorm's moment sums come from
    the GEMM epilogue, fp32 per 64 entries then fp64) against the same forward on precomputed edge features (a sweep over
    edge_attr in fp64): identical edge features, logits within 2e-6 * max|logit|, decisions identical outside that band; and
    against the fp64 oracle within the north-star tolerance.  Both graph builders (camera ids: layout known on the host; int64
    edge_index: decided on the device, the sweep is enqueued behind the flag and returns at once)."""
    params = mo.shipped_model_params(2, 1, 64, (48, 40))
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    sd = mo.init_weights(params, "resnet101", 5)
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    net.fuse_decisions = True
    for N, C_ in ((3000, 6), (700, 4)):
        x, ei, cam, _ = mo.synth_graph(N, C_, 9, D=64, planted=True)
        x[5] = x[4] + 1e-4 * torch.randn(64, generator=torch.Generator().manual_seed(1))       # a refined pair takes part in the sums
        ea = mo.edge_features(x, ei)
        ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
        outs = []
        for build in ("cameras", "edge_index", "precomputed"):
            if build == "cameras":
                d = Data(x=x.to(dev()), edge_index=None, mpn_graph=m.TrackletGraph.from_cameras(cam.numpy(), dev()), edge_attr=None)
            elif build == "edge_index":
                d = Data(x=x.to(dev()), edge_index=ei.to(dev()), edge_attr=None)
            else:
                d = Data(x=x.to(dev()), edge_index=ei.to(dev()), edge_attr=outs[0][2].clone())
            net._graphs.clear()
            out, h = net(d)
            torch.cuda.synchronize()
            outs.append((out["classified_edges"][-1].clone(), h.clone(), d.edge_attr.clone(), net.last_pred.clone()))
        scale = ref[-1].abs().max().item()
        assert torch.equal(outs[0][2], outs[1][2]) and torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][3], outs[1][3])
        assert (outs[0][0] - outs[2][0]).abs().max().item() <= 2e-6 * scale
        margin = (outs[2][0][:, 1] - outs[2][0][:, 0]).abs()
        assert not bool(((outs[0][3] != outs[2][3]) & (margin > 4e-6 * scale)).any())
        for o in outs:
            assert (o[0].cpu().double() - ref[-1]).abs().max().item() <= 1e-4 * scale
            assert (o[1].cpu().double() - href).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item())


class _NoComm:
    world, rank = 1, 0

    def all_reduce_sum(self, t):
        pass

    def all_gather_rows(self, full, blocks):
        pass

    def all_gather_ragged(self, t):
        return [t]


@pytest.mark.parametrize("L,n_cls,world", [(1, 1, 2), (3, 2, 3), (0, 1, 2)])
def test_sharded_cuda_shards_in_one_process(m, L, n_cls, world):
    """The N>1 path on one GPU: G row-block shards run phase by phase in one process (no concurrent waiting kernels)."""
    params = mo.shipped_model_params(L, n_cls, 64, (48, 40))
    x, ei, cam, _ = mo.synth_graph(90, 3, 41, D=64, planted=True)
    sd = mo.init_weights(params, "resnet101", 13)
    ea = mo.edge_features(x, ei)
    ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    outs, h, net = run_forward(m, params, sd, x, ei, ea)
    N, E = x.shape[0], ei.shape[1]
    rowptr = torch.searchsorted(ei[0].contiguous(), torch.arange(N + 1))
    blocks = m.partition_rows(rowptr, world)
    n_out = 1 if L == 0 else n_cls
    xd, W = x.to(dev()), net._weights(dev())
    shards, spans, keep = [], [], []
    for i, (n0, n1) in enumerate(blocks):
        lo, hi = m.shard_edges(ei, n0, n1)
        g = m.TrackletGraph(ei[:, lo:hi].to(dev()), N, row_offset=n0, n_rows=n1 - n0)
        logits = torch.empty(max(n_out, 1), hi - lo, 2, device=dev())
        ph = m.CudaPhases(g, W, xd, ea[lo:hi].to(dev()).contiguous(), L, n_cls, E, logits, None, None, False, ws_kind="shard%d" % i)
        shards.append(ph); spans.append((lo, hi)); keep.append((g, logits))
    m.sharded_forward(shards, _NoComm(), L, n_cls, blocks)
    torch.cuda.synchronize()
    for i in range(n_out):
        got = torch.cat([k[1][i] for k in keep]).cpu()
        assert (got.double() - ref[i]).abs().max().item() <= 1e-4 * ref[i].abs().max().item()
        assert (got - outs[i].cpu()).abs().max().item() <= 2e-6            # same kernels, only the reduction order differs
    hs = torch.cat([p.h_full()[b[0]:b[1]] for p, b in zip(shards, blocks)]).cpu()
    assert (hs.double() - href).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item())
    # sharded edge features: each shard computes its row block of the Gram matrix
    for (lo, hi), (g, _) in zip(spans, keep):
        ef = m.edge_features(xd, None, graph=g, use_tensor_cores=False).cpu()
        assert np.allclose(ef.numpy(), ea[lo:hi].numpy(), rtol=3e-6, atol=3e-6)
    for p in shards:
        p.close()


def _packed_batch(sizes, cams, D, seed):
    """PyG-style packing of several graphs: x concatenated, edge_index offset, ptr = node offsets."""
    xs, eis, ptr = [], [], [0]
    for i, (n, c) in enumerate(zip(sizes, cams)):
        x, ei, _, _ = mo.synth_graph(n, c, seed + i, D=D, planted=(i % 2 == 0))
        xs.append(x); eis.append(ei + ptr[-1]); ptr.append(ptr[-1] + n)
    return torch.cat(xs), torch.cat(eis, dim=1), torch.tensor(ptr, dtype=torch.int64), xs, eis


@pytest.mark.parametrize("D,tc", [(256, True), (2048, True), (48, True), (64, False)])
def test_batched_edge_features_block_diagonal(m, D, tc):
    sizes, cams = [40, 150, 30, 290, 52, 131], [4, 4, 3, 5, 4, 2]         # graphs larger than one 128-row tile included
    x, ei, ptr, xs, eis = _packed_batch(sizes, cams, D, 500)
    ref = mo.edge_features(x, ei, dtype=torch.float64).numpy()
    g = m.TrackletGraph(ei.to(dev()), x.shape[0], ptr=ptr.to(dev()))
    out = m.edge_features(x.to(dev()), None, graph=g, use_tensor_cores=tc).cpu().numpy()
    assert np.allclose(out, ref, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("L,n_cls,agg,reattach", [(1, 1, "sum", False), (3, 2, "sum", False), (2, 1, "max", False), (2, 1, "mean", False),
                                                  (3, 1, "sum", True)])
def test_batched_graphs_per_graph_batchnorm(m, L, n_cls, agg, reattach):
    """BASELINE configs[2]: many small graphs in one launch; statistics per graph == one reference forward per graph."""
    params = mo.shipped_model_params(L, n_cls, 64, (48, 40))
    params["node_agg_fn"] = agg
    params["reattach_initial_nodes"] = params["reattach_initial_edges"] = reattach
    sd = mo.init_weights(params, "resnet101", 17)
    sizes, cams = [40, 64, 30, 90, 52, 36], [4, 4, 3, 5, 4, 2]
    x, ei, ptr, xs, eis = _packed_batch(sizes, cams, 64, 300)
    ea = mo.edge_features(x, ei)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    net.fuse_decisions = True
    data = Data(x=x.to(dev()), edge_index=ei.to(dev()), edge_attr=ea.to(dev()), ptr=ptr.to(dev()))
    out, h = net(data)
    torch.cuda.synchronize()
    e0 = 0
    for gi, (xg, eig) in enumerate(zip(xs, eis)):
        n0, n1 = int(ptr[gi]), int(ptr[gi + 1])
        eg = eig.shape[1]
        ref, href = mo.mpn_forward(sd, params, "resnet101", xg, eig - n0, ea[e0:e0 + eg], dtype=torch.float64)
        for i in range(n_cls):
            got = out["classified_edges"][i][e0:e0 + eg].cpu().double()
            assert (got - ref[i]).abs().max().item() <= 1e-4 * ref[i].abs().max().item(), (gi, i)
        assert (h[n0:n1].cpu().double() - href).abs().max().item() <= 1e-4 * max(1.0, href.abs().max().item()), gi
        margin = (ref[-1][:, 1] - ref[-1][:, 0]).abs()
        bad = (net.last_pred[e0:e0 + eg].cpu().long() != ref[-1].argmax(1)) & (margin > 1e-4)
        assert not bool(bad.any())
        e0 += eg
    # a batch of one graph takes the single-graph path and must agree with it bit for bit
    d1 = Data(x=xs[0].to(dev()), edge_index=(eis[0] - 0).to(dev()), edge_attr=ea[:eis[0].shape[1]].to(dev()),
              ptr=torch.tensor([0, sizes[0]], device=dev()))
    d2 = Data(x=xs[0].to(dev()), edge_index=(eis[0] - 0).to(dev()), edge_attr=ea[:eis[0].shape[1]].to(dev()))
    o1, _ = net(d1); o2, _ = net(d2)
    assert torch.equal(o1["classified_edges"][-1], o2["classified_edges"][-1])
    with pytest.raises(ValueError):                                      # an edge between two graphs
        bad_ei = ei.clone(); bad_ei[1, 0] = int(ptr[2])
        m.TrackletGraph(bad_ei.to(dev()), x.shape[0], ptr=ptr.to(dev()))


def test_forward_rejects_cpu_and_training(m):
    params = mo.shipped_model_params(1, 1, 64, (48,))
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    x, ei, _, _ = mo.synth_graph(20, 2, 1, D=64)
    with pytest.raises(RuntimeError):
        net.eval()(Data(x=x, edge_index=ei, edge_attr=torch.zeros(ei.shape[1], 2)))
    net = net.to(dev()).train()
    with pytest.raises(NotImplementedError):
        net(Data(x=x.to(dev()), edge_index=ei.to(dev()), edge_attr=torch.zeros(ei.shape[1], 2, device=dev())))


def test_decide_kernel(m):
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(10001, 2, generator=g)
    logits[5] = torch.tensor([0.25, 0.25])                                # tie -> class 0
    prob, pred = mo.decide(logits)
    ld = logits.to(dev())
    p8 = torch.empty(10001, dtype=torch.uint8, device=dev())
    p1 = torch.empty(10001, device=dev())
    m._lib.check(m._lib.lib().mpn_decide(ld.data_ptr(), 10001, p8.data_ptr(), p1.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(p8.cpu().numpy(), pred.numpy().astype(np.uint8))
    assert (p1.cpu() - prob[:, 1]).abs().max().item() <= 1e-6


# ---------------------------------------------------------------------------------------------- post-processing
CONFIGS = (("full", (True, True, True)), ("cut_only", (True, False, False)), ("prune_only", (False, True, False)),
           ("split_only", (False, False, True)), ("cut_prune", (True, True, False)))


@pytest.mark.parametrize("path", POST_FILES, ids=ids(POST_FILES))
def test_post_processing_matches_reference_golden(m, path):
    g = np.load(path)
    N, Cn, _ = [int(v) for v in g["spec"]]
    ei = torch.from_numpy(np.stack([g["src"], g["dst"]]).astype(np.int64)).to(dev())
    data = Data(x=torch.zeros(N, 1, device=dev()), edge_index=ei)
    prob1 = torch.from_numpy(g["prob1"])
    preds_prob = torch.stack([1 - prob1, prob1], dim=1).to(dev())
    for tag, cfg in CONFIGS:
        pred = torch.from_numpy(g["pred"].astype(np.int64)).to(dev())
        CONFIG = {"CUTTING": cfg[0], "PRUNING": str(cfg[1]), "SPLITTING": cfg[2]}       # one flag as a CLI string
        ID, P = m.post_processing(Cn, None, None, pred, None, CONFIG, data, preds_prob)
        assert ID.dtype == torch.int64 and not ID.is_cuda and P.is_cuda and P.dtype == torch.int64
        assert np.array_equal(P.cpu().numpy(), g["pred_" + tag]), tag           # bit-exact decisions
        assert np.array_equal(ID.numpy(), g["labels_" + tag]), tag              # bit-exact label integers
        pred = torch.from_numpy(g["pred"].astype(np.int64)).to(dev())
        IDc, Pc = m.post_processing(Cn, None, None, pred, None, dict(CONFIG), data, preds_prob, numbering="canonical")
        assert np.array_equal(Pc.cpu().numpy(), g["pred_" + tag])
        assert same_partition(IDc.numpy(), g["labels_" + tag])
    lab0, _ = m.compute_SCC_and_Clusters(list(zip(g["src"][g["pred"] != 0].tolist(), g["dst"][g["pred"] != 0].tolist())), N)
    assert np.array_equal(lab0.numpy(), g["labels_initial"])


def test_stage_functions_mirror_reference_contracts(m):
    g = np.load(POST_FILES[0])
    N, Cn, _ = [int(v) for v in g["spec"]]
    src, dst = g["src"].astype(np.int64), g["dst"].astype(np.int64)
    ei = torch.from_numpy(np.stack([src, dst])).to(dev())
    data = Data(x=torch.zeros(N, 1, device=dev()), edge_index=ei)
    pred = torch.from_numpy(g["pred"].astype(np.int64)).to(dev())
    prob1 = torch.from_numpy(g["prob1"]).to(dev())
    new_pred, act_list = m.remove_edges_single_direction(None, pred, None, data=data)
    cut = po.cut_sequential(src, dst, g["pred"].astype(np.int64))
    assert np.array_equal(new_pred.cpu().numpy(), cut) and new_pred.data_ptr() != pred.data_ptr()
    assert act_list == [(int(src[e]), int(dst[e])) for e in np.flatnonzero(cut)]
    r = m.pruning(data, new_pred, prob1, None, Cn)
    ref = po.prune_sequential(src, dst, cut, g["prob1"], Cn, N)
    assert (r == [] and ref is None) or np.array_equal(r.cpu().numpy(), ref)
    assert m.pruning(data, torch.zeros_like(pred), prob1, None, Cn) == []            # nothing violated -> []
    p2 = pred.clone()
    out = m.splitting(None, p2, prob1, None, data, None, Cn)
    assert out.data_ptr() == p2.data_ptr()                                         # mutated in place
    assert np.array_equal(p2.cpu().numpy(), po.split_sequential(src, dst, g["pred"].astype(np.int64), g["prob1"], Cn, N))


@pytest.mark.parametrize("n_nodes,cams,seed", [(3000, 6, 1), (20000, 8, 2)])
def test_post_processing_sparse_large_vs_oracle_rounds(m, n_nodes, cams, seed):
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, seed, n_extra_per_node=6.0, flip_on=0.05,
                                                          flip_off=0.03, single_dir=0.05)
    lab_ref, act_ref = po.post_processing_rounds(src, dst, pred, prob, cams, n_nodes, numbering="reference")
    ei = torch.from_numpy(np.stack([src, dst])).to(dev())
    data = Data(x=torch.zeros(n_nodes, 1, device=dev()), edge_index=ei)
    p = torch.from_numpy(prob).to(dev())
    ID, P = m.post_processing(cams, None, None, torch.from_numpy(pred).to(dev()), None,
                              {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, data, p)
    assert np.array_equal(P.cpu().numpy(), act_ref)
    assert np.array_equal(ID.numpy(), lab_ref)
    assert np.bincount(ID.numpy()).max() <= cams                                    # size-independent property


def test_scc_with_one_directional_cycles(m):
    # 3-cycle of one-directional edges between mutual-edge pairs: SCC must merge them (condensation path)
    src = np.array([0, 1, 2, 3, 4, 5, 1, 3, 5, 6, 7], dtype=np.int64)
    dst = np.array([1, 0, 3, 2, 5, 4, 2, 4, 0, 7, 8], dtype=np.int64)
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    ei = torch.from_numpy(np.stack([src, dst])).to(dev())
    data = Data(x=torch.zeros(10, 1, device=dev()), edge_index=ei)
    pred = torch.ones(src.size, dtype=torch.int64, device=dev())
    prob = torch.linspace(0.6, 0.9, src.size).to(dev())
    ID, P = m.post_processing(8, None, None, pred, None, {"CUTTING": False, "PRUNING": True, "SPLITTING": True}, data, prob)
    ref, _ = po.scc_labels_reference(src, dst, np.ones(src.size), 10)
    assert np.array_equal(ID.numpy(), ref)
    IDc, _ = m.post_processing(8, None, None, pred, None, {"CUTTING": False, "PRUNING": True, "SPLITTING": True}, data, prob,
                               numbering="canonical")
    assert IDc.numpy().tolist() == [0, 0, 0, 0, 0, 0, 6, 7, 8, 9]


def _device_shards(m, src, dst, pred, prob1, n_nodes, world):
    ei = torch.from_numpy(np.stack([src, dst]).astype(np.int64))
    rowptr = torch.searchsorted(ei[0].contiguous(), torch.arange(n_nodes + 1))
    out = []
    for (n0, n1) in m.partition_rows(rowptr, world):
        lo, hi = m.shard_edges(ei, n0, n1)
        g = m.TrackletGraph(ei[:, lo:hi].to(dev()), n_nodes, row_offset=n0, n_rows=n1 - n0)
        out.append((g, torch.from_numpy(pred[lo:hi].astype(np.uint8)).to(dev()), torch.from_numpy(prob1[lo:hi]).to(dev())))
    return out


@pytest.mark.parametrize("path", POST_FILES, ids=ids(POST_FILES))
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_post_processing_matches_reference_golden(m, path, world):
    """Row-block shards (in one process): shard compaction, merged active list, rounds, write-back — decisions and label
    integers bit-exact against the reference's own post_processing outputs, for every flag combination."""
    g = np.load(path)
    N, Cn, _ = [int(v) for v in g["spec"]]
    for tag, cfg in CONFIGS:
        shards = _device_shards(m, g["src"], g["dst"], g["pred"], g["prob1"], N, world)
        CONFIG = {"CUTTING": cfg[0], "PRUNING": str(cfg[1]), "SPLITTING": cfg[2]}
        ID, preds = m.sharded_post_processing(Cn, shards, CONFIG, N, comm=_NoComm())
        assert np.array_equal(torch.cat(preds).cpu().numpy().astype(np.int64), g["pred_" + tag]), tag
        assert ID.dtype == torch.int64 and np.array_equal(ID.numpy(), g["labels_" + tag]), tag
    shards = _device_shards(m, g["src"], g["dst"], g["pred"], g["prob1"], N, world)       # no flag set: initial labels, no change
    ID, preds = m.sharded_post_processing(Cn, shards, {"CUTTING": False, "PRUNING": False, "SPLITTING": False}, N, comm=_NoComm())
    assert np.array_equal(ID.numpy(), g["labels_initial"]) and np.array_equal(torch.cat(preds).cpu().numpy(), g["pred"])


def test_sharded_post_processing_large_and_edge_cases(m):
    n_nodes, cams = 20000, 8
    src, dst, prob, pred, _ = po.planted_prediction_graph(n_nodes, cams, 5, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03,
                                                          single_dir=0.05)
    order = np.lexsort((dst, src))
    src, dst, prob, pred = src[order], dst[order], prob[order], pred[order]
    lab_ref, act_ref = po.post_processing_rounds(src, dst, pred, prob, cams, n_nodes, numbering="reference")
    shards = _device_shards(m, src, dst, pred, prob, n_nodes, 4)
    single = m.sharded_post_processing(cams, shards[0], {"CUTTING": False, "PRUNING": False, "SPLITTING": False}, n_nodes,
                                       comm=_NoComm())                                     # a bare triple -> a bare tensor back
    assert isinstance(single[1], torch.Tensor) and single[1].data_ptr() == shards[0][1].data_ptr()
    ID, preds = m.sharded_post_processing(cams, shards, {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, n_nodes, comm=_NoComm())
    assert np.array_equal(torch.cat(preds).cpu().numpy().astype(np.int64), act_ref) and np.array_equal(ID.numpy(), lab_ref)
    assert np.bincount(ID.numpy()).max() <= cams
    IDc, _ = m.sharded_post_processing(cams, _device_shards(m, src, dst, pred, prob, n_nodes, 4),
                                       {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, n_nodes, comm=_NoComm(),
                                       numbering="canonical")
    assert same_partition(IDc.numpy(), lab_ref)
    # nothing active anywhere: every node its own cluster, decisions untouched
    shards = _device_shards(m, src, dst, 0 * pred, prob, n_nodes, 2)
    ID, preds = m.sharded_post_processing(cams, shards, {"CUTTING": True, "PRUNING": True, "SPLITTING": True}, n_nodes, comm=_NoComm())
    assert ID.tolist() == list(range(n_nodes)) and not any(bool(p.any()) for p in preds)
    with pytest.raises(ValueError):
        m.sharded_post_processing(cams, (shards[0][0], shards[0][1].long(), shards[0][2]), {"CUTTING": True, "PRUNING": True,
                                                                                              "SPLITTING": True}, n_nodes, comm=_NoComm())
    with pytest.raises(RuntimeError):
        m.sharded_post_processing(cams, (shards[0][0], shards[0][1].cpu(), shards[0][2]), {"CUTTING": True, "PRUNING": True,
                                                                                             "SPLITTING": True}, n_nodes, comm=_NoComm())


def test_multi_gpu_fused_collectives_torchrun(m):
    """N>1 on real GPUs (skipped on a single-GPU box): fused peer-memory collectives and the NCCL schedule vs the oracle."""
    import os
    import subprocess
    import sys
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 4)),
           "--master-addr", "127.0.0.1", "--master-port", "29577", os.path.join(root, "tests", "dist_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0 and "dist_gpu_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_graph_stream_packed_decisions(m):
    """GraphStream(packed_decisions=True): the bit mask that crosses PCIe unpacks to exactly the per-edge decisions (ragged E: not
    a multiple of 32), eager and replayed as a CUDA graph; mpn_pack_decisions on its own for E = 1, 31, 32, 33."""
    d0 = dev()
    lib = m._lib.lib()
    for E in (1, 31, 32, 33, 1000):
        pred = (torch.rand(E, device=d0) < 0.3).to(torch.uint8)
        words = torch.zeros(4 * ((E + 31) // 32), dtype=torch.uint8, device=d0)
        m._lib.check(lib.mpn_pack_decisions(pred.data_ptr(), E, words.data_ptr(), torch.cuda.current_stream(d0).cuda_stream))
        assert np.array_equal(m.unpack_decisions(words.cpu(), E), pred.cpu().numpy())
    params = mo.shipped_model_params(1, 1, 64, (32,))
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101").to(d0).eval()
    n, cams = 45, 3
    cam = (np.arange(n) * cams // n)
    x = torch.randn(n, 64).pin_memory()
    ref_stream = m.GraphStream(net, d0, depth=1)
    E = int(sum((cam != c).sum() for c in cam))
    want = torch.empty(E, dtype=torch.uint8).pin_memory()
    ref_stream.submit(x, cam, want)
    ref_stream.drain()
    for replay in (False, True):
        gs = m.GraphStream(net, d0, depth=2, graph_replay=replay, packed_decisions=True)
        bufs = [torch.zeros(4 * ((E + 31) // 32), dtype=torch.uint8).pin_memory() for _ in range(3)]
        for i in range(5):
            gs.submit(x, cam, bufs[i % 3])
        gs.drain()
        for bts in bufs:
            assert np.array_equal(m.unpack_decisions(bts, E), want.numpy())
        with pytest.raises(ValueError):
            gs.submit(x, cam, want)                       # a per-edge buffer where the bit mask is expected


def test_more_class_steps_than_enc_steps(m):
    """num_class_steps > num_enc_steps: first_class_step <= 0, the reference classifies at every step and returns L outputs
    (models/mpn.py:281,290); so does this forward."""
    params = mo.shipped_model_params(2, 3, 64, (32,))
    x, ei, cam, _ = mo.synth_graph(60, 3, 5, D=64)
    sd = mo.init_weights(params, "resnet101", 4)
    ea = mo.edge_features(x, ei)
    ref, href = mo.mpn_forward(sd, params, "resnet101", x, ei, ea, dtype=torch.float64)
    assert len(ref) == 2
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(sd, strict=True)
    net = net.to(dev()).eval()
    out, h = net(Data(x=x.to(dev()), edge_index=ei.to(dev()), edge_attr=ea.to(dev())))
    assert len(out["classified_edges"]) == 2
    for got, want in zip(out["classified_edges"], ref):
        scale = max(1.0, want.abs().max().item())
        assert (got.cpu().double() - want).abs().max().item() <= 2e-5 * scale
