"""sys.modules shims that let the UNMODIFIED reference (``/root/reference``) import in the build container.

Used only by ``tests/golden/make_golden.py`` (fixture generation) and, when ``/root/reference`` exists,
by ``tests/test_reference_live.py``.  Nothing here is on the product path and nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` reads ``/root/reference``.

The reference needs ``torch_scatter`` (models/mpn.py:4, utils.py:20), ``matplotlib`` (utils.py:19) and
``torch_geometric`` (inference.py:13) which are not installed; the shims restate only the call
signatures the hot path uses.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MPN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "mpn.py"))


def install():
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")

        def scatter_add(src, index, dim=0, dim_size=None, out=None):
            shape = list(src.shape)
            shape[dim] = dim_size if dim_size is not None else int(index.max()) + 1
            return torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(dim, index, src)

        # torch_scatter 2.0.8 (env_gnn.yml:98; not installable here) documented semantics, dim=0 + dim_size as mpn.py:196-202 calls:
        def scatter_mean(src, index, dim=0, dim_size=None, out=None):
            total = scatter_add(src, index, dim, dim_size)
            cnt = torch.zeros(total.shape[dim], dtype=src.dtype, device=src.device).index_add_(
                0, index, torch.ones(index.numel(), dtype=src.dtype, device=src.device))
            return total / cnt.clamp_(min=1).view([-1] + [1] * (src.dim() - 1))        # count clamped to >= 1

        def scatter_max(src, index, dim=0, dim_size=None, out=None):
            shape = list(src.shape)
            shape[dim] = dim_size if dim_size is not None else int(index.max()) + 1
            res = torch.full(shape, torch.finfo(src.dtype).min, dtype=src.dtype, device=src.device)
            idx = index.view([-1] + [1] * (src.dim() - 1)).expand_as(src)
            res = res.scatter_reduce(dim, idx, src, reduce="amax", include_self=True)
            res = res.masked_fill(res == torch.finfo(src.dtype).min, 0)                 # untouched entries -> 0
            return res, None                                                            # (values, argmax): mpn.py:199 takes [0]

        ts.scatter_add, ts.scatter_mean, ts.scatter_max = scatter_add, scatter_mean, scatter_max
        sys.modules["torch_scatter"] = ts
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].use = lambda *a, **k: None
    if "torch_geometric" not in sys.modules:
        tg, tgd, tgu = (types.ModuleType(n) for n in ("torch_geometric", "torch_geometric.data", "torch_geometric.utils"))

        class Data:
            def __init__(self, **kw):
                self.__dict__.update(kw)

            @property
            def num_nodes(self):
                return self.x.size(0)

        tgd.Data, tgd.Batch, tgu.to_networkx = Data, object, None
        tg.data, tg.utils = tgd, tgu
        sys.modules.update({"torch_geometric": tg, "torch_geometric.data": tgd, "torch_geometric.utils": tgu})
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_reference():
    """Returns (MOTMPNet, utils module, inference module, Data class) of the unmodified reference."""
    install()
    from models.mpn import MOTMPNet          # noqa: E402
    import utils as ref_utils                # noqa: E402
    import inference as ref_inference        # noqa: E402
    return MOTMPNet, ref_utils, ref_inference, sys.modules["torch_geometric.data"].Data
