"""BASELINE configs[3] post-processing once more, with the library's stage timing on stderr (MPN_POST_DEBUG=1):

    MPN_POST_DEBUG=1 python tools/post_profile.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gcn_mtmc_b200 as m

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
m._lib.require_device(0)
print(json.dumps(bench.extra_post_processing(m, dev), indent=1))
