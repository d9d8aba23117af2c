"""CUDA post-processing (through the C ABI) against the plain-C oracle at a size the Python restatement needs minutes for.
B200 only.  Named to run last: it was added after the round's GPU time was spent, and the same input is checked on the CPU
in tests/test_c_oracle.py (C oracle == Python rounds oracle, same generator arguments)."""
import numpy as np
import pytest
import torch

from oracle import postproc_c as pc
from oracle import postproc_oracle as po
from tests._util import Data, same_partition

pytestmark = pytest.mark.gpu

# the same arguments as tests/test_c_oracle.py::test_c_oracle_pins_the_gpu_case
GPU_CASE = dict(n_nodes=200000, cams=8, seed=3, n_extra_per_node=6.0, flip_on=0.05, flip_off=0.03, single_dir=0.05)


@pytest.fixture(scope="module")
def m():
    import gcn_mtmc_b200 as mod
    mod._lib.require_device(0)
    return mod


def test_post_processing_200k_nodes_vs_c_oracle(m):
    c = GPU_CASE
    dev = torch.device("cuda", 0)
    src, dst, prob, pred, _ = po.planted_prediction_graph(c["n_nodes"], c["cams"], c["seed"], n_extra_per_node=c["n_extra_per_node"],
                                                          flip_on=c["flip_on"], flip_off=c["flip_off"], single_dir=c["single_dir"])
    lab_ref, act_ref = pc.post_processing(src, dst, pred, prob, c["cams"], c["n_nodes"], numbering="reference")
    data = Data(x=torch.zeros(c["n_nodes"], 1, device=dev), edge_index=torch.from_numpy(np.stack([src, dst])).to(dev))
    cfg = {"CUTTING": True, "PRUNING": True, "SPLITTING": True}
    ID, P = m.post_processing(c["cams"], None, None, torch.from_numpy(pred).to(dev), None, cfg, data, torch.from_numpy(prob).to(dev))
    assert np.array_equal(P.cpu().numpy(), act_ref)                                # decisions: bit-exact
    assert np.array_equal(ID.numpy(), lab_ref)                                     # the reference's label integers: bit-exact
    assert np.bincount(ID.numpy()).max() <= c["cams"]                              # size-independent property
    IDc, Pc = m.post_processing(c["cams"], None, None, torch.from_numpy(pred).to(dev), None, cfg, data, torch.from_numpy(prob).to(dev),
                                numbering="canonical")
    assert np.array_equal(Pc.cpu().numpy(), act_ref) and same_partition(IDc.numpy(), lab_ref)
