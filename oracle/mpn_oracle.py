"""CPU oracle for the tracklet-graph message-passing path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-torch (CPU) restatement of the reference algorithm; it is the
checker, never the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity status: the reference ships no golden vectors, weights or tests for this path
("parity unpinned" by the reference's own tests).  The restatement is instead pinned
against outputs of the reference code itself, run in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` replays them).

Reference lines followed (relative to the upstream repository root):
  * MLP block ``Linear -> BatchNorm1d(track_running_stats=False) -> ReLU -> Dropout``:
    ``models/mlp.py:4-33``  (batch statistics are used even in ``eval()``).
  * encoder / classifier ``MLPGraphIndependent``: ``models/mpn.py:103-142``.
  * edge update ``EdgeModel.forward``: ``models/mpn.py:67-69``.
  * node update ``NodeModel.forward``: ``models/mpn.py:97-99`` with ``scatter_add``
    (``models/mpn.py:202``).
  * step loop / classification schedule ``MOTMPNet.forward``: ``models/mpn.py:250-299``.
  * edge features: ``inference.py:453-456``; column-wise normalisation ``inference.py:403-404``.
  * decisions: ``inference.py:475-479``.
  * shipped hyper-parameters: ``config/config_training.yaml:68-111``.
"""
from __future__ import annotations

import copy
import math
from collections import OrderedDict

import torch

BN_EPS = 1e-5           # nn.BatchNorm1d default, models/mlp.py:16
PAIRWISE_EPS = 1e-6     # F.pairwise_distance default eps, inference.py:453
COSINE_EPS = 1e-8       # F.cosine_similarity default eps, inference.py:454


def shipped_model_params(num_enc_steps: int = 1, num_class_steps: int = 1,
                         node_in_dim: int = 2048, node_fc_dims=(1024, 512, 128)) -> dict:
    """GRAPH_NET_PARAMS as shipped (config/config_training.yaml:68-111), restated."""
    return {
        "node_agg_fn": "sum",
        "num_enc_steps": num_enc_steps,
        "num_class_steps": num_class_steps,
        "reattach_initial_nodes": False,
        "reattach_initial_edges": False,
        "encoder_feats_dict": {
            "edges": {"edge_in_dim": 2, "edge_fc_dims": [4], "edge_out_dim": 4},
            "nodes": {"resnet101": {"node_in_dim": node_in_dim, "node_fc_dims": list(node_fc_dims),
                                    "node_out_dim": 32, "dropout_p": 0.1, "use_batchnorm": True}},
        },
        "edge_model_feats_dict": {"fc_dims": [4], "dropout_p": 0.1, "use_batchnorm": True},
        "node_model_feats_dict": {"fc_dims": [32], "dropout_p": 0.1, "use_batchnorm": True},
        "classifier_feats_dict": {"edge_in_dim": 4, "edge_fc_dims": [], "edge_out_dim": 2,
                                  "dropout_p": 0, "use_batchnorm": False, "is_classifier": True},
    }


# --------------------------------------------------------------------------------------
# state_dict layout (models/mlp.py:11-30): position of every module inside fc_layers
# --------------------------------------------------------------------------------------
def mlp_layout(input_dim, fc_dims, dropout_p, use_batchnorm, is_classifier=False):
    """Return [(linear_idx, bn_idx_or_None, relu?, in_dim, out_dim), ...] as nn.Sequential numbers them."""
    out, idx = [], 0
    for dim in fc_dims:
        lin = idx
        idx += 1
        bn = None
        relu = False
        if not is_classifier:
            if use_batchnorm and dim != 1:
                bn = idx
                idx += 1
            if dim != 1:
                relu = True
                idx += 1
            if dropout_p is not None and dim != 1:
                idx += 1
        out.append((lin, bn, relu, input_dim, dim))
        input_dim = dim
    return out


def model_layouts(model_params: dict, arch: str) -> "OrderedDict[str, list]":
    """prefix -> layout for every MLP of MOTMPNet (models/mpn.py:166-247)."""
    p = copy.deepcopy(model_params)
    enc = dict(p["encoder_feats_dict"]["edges"])
    enc.update(p["encoder_feats_dict"]["nodes"][arch])
    cls = p["classifier_feats_dict"]
    edge_factor = 2 if p["reattach_initial_edges"] else 1
    node_factor = 2 if p["reattach_initial_nodes"] else 1
    edge_in = node_factor * 2 * enc["node_out_dim"] + edge_factor * enc["edge_out_dim"]
    node_in = node_factor * enc["node_out_dim"] + enc["edge_out_dim"]
    em, nm = p["edge_model_feats_dict"], p["node_model_feats_dict"]
    lay = OrderedDict()
    lay["encoder.node_mlp.fc_layers"] = mlp_layout(enc["node_in_dim"], list(enc["node_fc_dims"]) + [enc["node_out_dim"]],
                                                   enc["dropout_p"], enc["use_batchnorm"])
    lay["encoder.edge_mlp.fc_layers"] = mlp_layout(enc["edge_in_dim"], list(enc["edge_fc_dims"]) + [enc["edge_out_dim"]],
                                                   enc["dropout_p"], enc["use_batchnorm"])
    lay["classifier.edge_mlp.fc_layers"] = mlp_layout(cls["edge_in_dim"], list(cls["edge_fc_dims"]) + [cls["edge_out_dim"]],
                                                      cls.get("dropout_p"), cls.get("use_batchnorm"),
                                                      cls.get("is_classifier", False))
    lay["MPNet.edge_model.edge_mlp.fc_layers"] = mlp_layout(edge_in, em["fc_dims"], em["dropout_p"], em["use_batchnorm"])
    lay["MPNet.node_model.node_mlp.fc_layers"] = mlp_layout(node_in, nm["fc_dims"], nm["dropout_p"], nm["use_batchnorm"])
    return lay


def init_weights(model_params: dict, arch: str = "resnet101", seed: int = 0,
                 affine_jitter: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic random-init state_dict with the reference's key names and shapes.

    nn.Linear-style U(-1/sqrt(in), 1/sqrt(in)) for weight and bias; BatchNorm affine either the
    default (1, 0) or jittered (weight~U(0.5,1.5), bias~N(0,0.1)) so that BN is not trivially
    identity-affine.  Independent of the reference constructor's RNG use.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for prefix, lay in model_layouts(model_params, arch).items():
        for lin, bn, _relu, din, dout in lay:
            bound = 1.0 / math.sqrt(din)
            sd[f"{prefix}.{lin}.weight"] = (torch.rand(dout, din, generator=g) * 2 - 1) * bound
            sd[f"{prefix}.{lin}.bias"] = (torch.rand(dout, generator=g) * 2 - 1) * bound
            if bn is not None:
                if affine_jitter:
                    sd[f"{prefix}.{bn}.weight"] = torch.rand(dout, generator=g) + 0.5
                    sd[f"{prefix}.{bn}.bias"] = torch.randn(dout, generator=g) * 0.1
                else:
                    sd[f"{prefix}.{bn}.weight"] = torch.ones(dout)
                    sd[f"{prefix}.{bn}.bias"] = torch.zeros(dout)
    return sd


# --------------------------------------------------------------------------------------
# forward pieces
# --------------------------------------------------------------------------------------
def batchnorm_batchstats(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """BatchNorm1d(track_running_stats=False): biased batch variance, eps 1e-5 (models/mlp.py:16)."""
    if x.shape[0] <= 1:
        raise ValueError("Expected more than 1 value per channel when training, got input size %s" % (list(x.shape),))
    mean = x.mean(dim=0)
    var = x.var(dim=0, unbiased=False)
    return (x - mean) / torch.sqrt(var + BN_EPS) * weight + bias


def mlp_forward(sd, prefix: str, layout, x: torch.Tensor) -> torch.Tensor:
    """models/mlp.py:32-33 over the layout of models/mlp.py:11-30 (Dropout is identity in eval)."""
    for lin, bn, relu, _din, _dout in layout:
        w = sd[f"{prefix}.{lin}.weight"].to(x.dtype)
        b = sd[f"{prefix}.{lin}.bias"].to(x.dtype)
        x = x @ w.t() + b
        if bn is not None:
            x = batchnorm_batchstats(x, sd[f"{prefix}.{bn}.weight"].to(x.dtype), sd[f"{prefix}.{bn}.bias"].to(x.dtype))
        if relu:
            x = torch.relu(x)
    return x


def scatter_add_rows(src: torch.Tensor, index: torch.Tensor, dim_size: int) -> torch.Tensor:
    """torch_scatter.scatter_add(src, index, dim=0, dim_size) (models/mpn.py:202): sequential on CPU."""
    return torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)


def aggregate_rows(src: torch.Tensor, index: torch.Tensor, dim_size: int, how: str) -> torch.Tensor:
    """node_agg_fn of models/mpn.py:193-202 = torch_scatter 2.0.8 (env_gnn.yml:98, not vendored) scatter_add / scatter_mean /
    scatter_max along dim 0 with dim_size: mean divides the sum by the per-row count clamped to >= 1; max of a row without
    elements is 0 (the library fills untouched entries of a fresh output with 0)."""
    if how == "sum":
        return scatter_add_rows(src, index, dim_size)
    if how == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device).index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype, device=src.device))
        return scatter_add_rows(src, index, dim_size) / cnt.clamp_min(1).unsqueeze(1)
    if how == "max":
        out = torch.full((dim_size,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
        out = out.scatter_reduce(0, index.unsqueeze(1).expand_as(src), src, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out), torch.zeros_like(out), out)
    raise ValueError(how)


def mpn_forward(sd, model_params: dict, arch: str, x: torch.Tensor, edge_index: torch.Tensor,
                edge_attr: torch.Tensor, dtype=torch.float32):
    """MOTMPNet.forward (models/mpn.py:250-299).  Returns (list of [E,2] logits, h [N,node_out]).  Plain torch ops: runs on
    whatever device ``x`` lives on (the fp64 checks at benchmark size run it on the GPU; ``sd`` is moved along)."""
    lay = model_layouts(model_params, arch)
    if any(v.device != x.device for v in sd.values()):
        sd = {k: v.to(x.device) for k, v in sd.items()}
    x = x.to(dtype)
    edge_attr = edge_attr.to(dtype)
    row, col = edge_index[0].long(), edge_index[1].long()
    L = int(model_params["num_enc_steps"])
    n_cls = int(model_params["num_class_steps"])
    re_n, re_e = model_params["reattach_initial_nodes"], model_params["reattach_initial_edges"]

    # encoder (models/mpn.py:270): note the (edge_feats, nodes_feats) argument order at mpn.py:128
    e = mlp_forward(sd, "encoder.edge_mlp.fc_layers", lay["encoder.edge_mlp.fc_layers"], edge_attr)
    h = mlp_forward(sd, "encoder.node_mlp.fc_layers", lay["encoder.node_mlp.fc_layers"], x)
    e0, h0 = e, h
    outs = []
    first_class_step = L - n_cls + 1
    for step in range(1, L + 1):
        if re_e:
            e = torch.cat((e0, e), dim=1)
        if re_n:
            h = torch.cat((h0, h), dim=1)
        # edge update (mpn.py:48,68-69)
        e = mlp_forward(sd, "MPNet.edge_model.edge_mlp.fc_layers", lay["MPNet.edge_model.edge_mlp.fc_layers"],
                        torch.cat([h[row], h[col], e], dim=1))
        # node update (mpn.py:97-99): messages use the node itself (x[row]) and the edge feature
        m = mlp_forward(sd, "MPNet.node_model.node_mlp.fc_layers", lay["MPNet.node_model.node_mlp.fc_layers"],
                        torch.cat([h[row], e], dim=1))
        h = aggregate_rows(m, row, h.shape[0], model_params.get("node_agg_fn", "sum"))
        if step >= first_class_step:
            outs.append(mlp_forward(sd, "classifier.edge_mlp.fc_layers", lay["classifier.edge_mlp.fc_layers"], e))
    if L == 0:
        outs.append(mlp_forward(sd, "classifier.edge_mlp.fc_layers", lay["classifier.edge_mlp.fc_layers"], e))
    return outs, h


def edge_features(x: torch.Tensor, edge_index: torch.Tensor, chunk: int = 1 << 17, dtype=torch.float32) -> torch.Tensor:
    """edge_attr = [ ||x_r - x_c + 1e-6||_2 , 1 - cos(x_r, x_c) ]  (inference.py:453-456).

    pairwise_distance adds eps to the difference before the norm; cosine_similarity divides the dot
    product by max(||a||*||b||, 1e-8).  Chunked so the two [E,D] gathers never materialise at once.
    """
    x = x.to(dtype)
    row, col = edge_index[0].long(), edge_index[1].long()
    E = row.numel()
    out = torch.empty(E, 2, dtype=dtype, device=x.device)
    for s in range(0, E, chunk):
        a, b = x[row[s:s + chunk]], x[col[s:s + chunk]]
        out[s:s + chunk, 0] = (a - b + PAIRWISE_EPS).norm(dim=1)
        denom = (a.norm(dim=1) * b.norm(dim=1)).clamp_min(COSINE_EPS)
        out[s:s + chunk, 1] = 1 - (a * b).sum(dim=1) / denom
    return out


def decide(logits: torch.Tensor):
    """softmax(dim=1) and argmax(dim=1) (inference.py:475-479); ties resolve to class 0."""
    prob = torch.softmax(logits, dim=1)
    pred = torch.argmax(logits, dim=1)
    return prob, pred


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def cross_camera_edge_index(cam: torch.Tensor) -> torch.Tensor:
    """For each camera ascending, cartesian_prod(in_cam, out_cam), concatenated (inference.py:407-413)."""
    nodes = torch.arange(cam.numel())
    parts = []
    for c in torch.unique(cam).tolist():
        parts.append(torch.cartesian_prod(nodes[cam == c], nodes[cam != c]))
    return torch.cat(parts, dim=0).t().contiguous()


def thin_edges(edge_index: torch.Tensor, seed: int, keep: float = 0.7, empty_rows=(3, 7)) -> torch.Tensor:
    """Drop a random 30% of the edges and every edge leaving ``empty_rows`` (rows that then aggregate nothing)."""
    g = torch.Generator().manual_seed(1000 + seed)
    m = torch.rand(edge_index.shape[1], generator=g) < keep
    for r in empty_rows:
        m &= edge_index[0] != r
    return edge_index[:, m].contiguous()


def balanced_cameras(N: int, C: int) -> torch.Tensor:
    return (torch.arange(N) * C // N).to(torch.int64)


def synth_graph(N: int, C: int, seed: int, D: int = 2048, planted: bool = False, noise: float = 0.3):
    """Graph(N,C,seed): nodes sorted by camera, column-normalised features, dense cross-camera edges."""
    g = torch.Generator().manual_seed(seed)
    cam = balanced_cameras(N, C)
    if planted:
        n_id = max(2, int(N / (C * 0.7)))
        cent = torch.randn(n_id, D, generator=g)
        ident = torch.randint(0, n_id, (N,), generator=g)
        x = cent[ident] + noise * torch.randn(N, D, generator=g)
    else:
        ident = torch.arange(N)
        x = torch.randn(N, D, generator=g)
    x = torch.nn.functional.normalize(x, p=2, dim=0)          # inference.py:403-404 (per column!)
    return x, cross_camera_edge_index(cam), cam, ident
