"""The reference ITSELF as the CPU arm.  TEST / BENCH INFRASTRUCTURE ONLY (never imported by the product).

``build_ref()`` (called by ``__graft_entry__.build()`` when ``/root/reference`` exists, i.e. in the build container) copies the
four Python files of the hot path and the shipped config VERBATIM from the read-only reference tree into ``oracle/_ref/``:

    models/mpn.py  models/mlp.py  utils.py  inference.py  config/config_training.yaml

``oracle/_ref/`` is git-ignored (no reference source enters the history) but not gpurun-ignored, so the files travel to the
GPU box, where ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg run them on the host cores.  The three
third-party modules the reference imports and this image lacks (torch_scatter, torch_geometric, matplotlib) are replaced by
the same ~40 lines of ``sys.modules`` shims that generated the golden vectors (tests/golden/ref_shims.py, SURVEY.md
appendix A); nothing of the reference's own code is touched.

What is timed:
  * ``MOTMPNet.forward(data)``                models/mpn.py:250-299   — called as inference.py:469 calls it;
  * the edge-feature statements               inference.py:453-456   — they are four lines in the middle of a 200-line driver
    function, not a callable: the statements are read from the copied file at run time (located by their text, the line numbers
    are checked) and executed unchanged on consecutive 200 k-edge chunks of the edge list (un-chunked they need 2 x E x 8 KB);
  * softmax / argmax                          inference.py:475-479.
"""
import ast
import copy
import os
import shutil
import sys
import textwrap
import time

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
FILES = ("models/mpn.py", "models/mlp.py", "utils.py", "inference.py", "config/config_training.yaml")
EF_FIRST_LINE, EF_LAST_LINE = 453, 456            # inference.py: node_dist_g = ... ; edge_attr = torch.cat(...)


def build_ref(reference_root: str = "/root/reference") -> bool:
    """Copy the hot-path files of the reference into oracle/_ref (verbatim).  False when the reference tree is absent."""
    if not os.path.isfile(os.path.join(reference_root, "models", "mpn.py")):
        return False
    for rel in FILES:
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference_root, rel), dst)
    open(os.path.join(REF_DIR, "models", "__init__.py"), "a").close()
    return True


def available() -> bool:
    return all(os.path.isfile(os.path.join(REF_DIR, rel)) for rel in FILES)


_loaded = None


def load():
    """(MOTMPNet class, GRAPH_NET_PARAMS dict, compiled edge-feature statements, Data class) of the copied reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("oracle/_ref is not built (python -c 'import __graft_entry__ as g; g.build()' in the build container)")
    sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "tests", "golden"))
    import ref_shims
    ref_shims.REFERENCE_ROOT = REF_DIR
    ref_shims.install()                               # puts REF_DIR on sys.path after installing the three module shims
    from models.mpn import MOTMPNet                   # noqa: E402  (oracle/_ref/models/mpn.py)
    import yaml
    with open(os.path.join(REF_DIR, "config", "config_training.yaml")) as f:
        params = yaml.safe_load(f)["GRAPH_NET_PARAMS"]
    with open(os.path.join(REF_DIR, "inference.py")) as f:
        lines = f.read().split("\n")
    stmts = textwrap.dedent("\n".join(lines[EF_FIRST_LINE - 1:EF_LAST_LINE]))
    if "F.pairwise_distance" not in stmts or "F.cosine_similarity" not in stmts or "edge_attr = torch.cat" not in stmts:
        raise RuntimeError("inference.py:%d-%d of oracle/_ref are not the edge-feature statements" % (EF_FIRST_LINE, EF_LAST_LINE))
    ast.parse(stmts)
    code = compile(stmts, os.path.join(REF_DIR, "inference.py") + ":%d-%d" % (EF_FIRST_LINE, EF_LAST_LINE), "exec")
    _loaded = (MOTMPNet, params, code, sys.modules["torch_geometric.data"].Data)
    return _loaded


def make_model(L: int = 1, n_cls: int = 1, seed: int = 0):
    """The reference module with its default initialisation under torch.manual_seed(seed)."""
    MOTMPNet, params, _, _ = load()
    p = copy.deepcopy(params)                         # the constructor mutates its argument (models/mpn.py:167-170)
    p["num_enc_steps"], p["num_class_steps"] = L, n_cls
    torch.manual_seed(seed)
    return MOTMPNet(p, None, "resnet101").eval()


def edge_features(x, edge_index, chunk: int = 200_000):
    """inference.py:453-456 executed verbatim on consecutive chunks of the edge list -> edge_attr [E,2]."""
    import torch.nn.functional as F
    _, _, code, _ = load()
    out = torch.empty(edge_index.shape[1], 2)
    for s in range(0, edge_index.shape[1], chunk):
        env = {"F": F, "torch": torch, "node_embeds_g": x, "edge_ixs_g": edge_index[:, s:s + chunk]}
        exec(code, env)
        out[s:s + chunk] = env["edge_attr"]
    return out


def timed_step(model, x, edge_index, ef_edges: int, edge_attr=None):
    """One bounded sample of the configs[1] step on the host cores: the reference forward on the WHOLE graph, its edge-feature
    statements on the first ``ef_edges`` edges (rate per edge), softmax / argmax.  Returns seconds per piece and the outputs."""
    import torch.nn.functional as F          # noqa: F401
    _, _, _, Data = load()
    E = edge_index.shape[1]
    n_ef = int(min(E, max(1, ef_edges)))
    t0 = time.perf_counter()
    ea_part = edge_features(x, edge_index[:, :n_ef])
    t_ef = time.perf_counter() - t0
    if edge_attr is None:
        edge_attr = ea_part if n_ef == E else None
    if edge_attr is None:
        raise ValueError("edge_attr of the whole graph is needed when only a slice of the edge features is timed")
    data = Data(x=x, edge_index=edge_index, edge_attr=edge_attr)
    with torch.no_grad():
        t0 = time.perf_counter()
        outputs, h = model(data)                      # inference.py:469
        t_fwd = time.perf_counter() - t0
        t0 = time.perf_counter()
        preds = outputs['classified_edges'][-1]
        preds_prob = torch.nn.Softmax(dim=1)(preds)   # inference.py:475-477
        predictions = torch.argmax(preds, dim=1)      # inference.py:479
        t_dec = time.perf_counter() - t0
    return {"t_forward": t_fwd, "t_edge_features": t_ef, "ef_edges": n_ef, "t_decisions": t_dec, "E": E,
            "logits": preds, "prob": preds_prob, "pred": predictions, "h": h, "edge_attr_slice": ea_part}
