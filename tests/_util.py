import glob
import os

import numpy as np
import torch

from oracle import mpn_oracle as mo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MPN_FILES = sorted(glob.glob(os.path.join(GOLDEN, "mpn_*.npz")))
POST_FILES = sorted(glob.glob(os.path.join(GOLDEN, "post_*.npz")))


class Data:
    """Stand-in for torch_geometric.data.Data (inference.py:458): attribute bag with num_nodes."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def num_nodes(self):
        return self.x.size(0)


def ids(files):
    return [os.path.basename(p)[:-4] for p in files]


def load_mpn_case(path):
    g = np.load(path)
    spec = [int(v) for v in g["spec"]]
    N, C, gseed, wseed, L, n_cls, din, planted, jitter = spec[:9]
    agg, thin = (("sum", "mean", "max")[spec[9]], bool(spec[10])) if len(spec) > 9 else ("sum", False)
    fcd = tuple(int(v) for v in g["fc_dims"])
    params = mo.shipped_model_params(L, n_cls, din, fcd)
    params["node_agg_fn"] = agg
    if len(spec) > 11:
        params["reattach_initial_nodes"], params["reattach_initial_edges"] = bool(spec[11]), bool(spec[12])
    x, edge_index, cam, _ = mo.synth_graph(N, C, gseed, D=din, planted=bool(planted))
    if thin:
        edge_index = mo.thin_edges(edge_index, gseed)
    assert np.array_equal(edge_index.numpy(), g["edge_index"].astype(np.int64)), "generator drift: edge_index"
    assert np.allclose([x.double().sum().item(), x.double().abs().sum().item()], g["x_checksum"], rtol=1e-12)
    sd = mo.init_weights(params, "resnet101", wseed, affine_jitter=bool(jitter))
    assert np.allclose(sum(v.double().sum().item() for v in sd.values()), g["w_checksum"][0], rtol=1e-12)
    return g, params, sd, x, edge_index, C


def same_partition(a, b):
    a, b = np.asarray(a).tolist(), np.asarray(b).tolist()
    return len(set(zip(a, b))) == len(set(a)) == len(set(b))
