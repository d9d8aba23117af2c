"""B200-native tracklet-graph message passing (drop-in for models/mpn.py + the post-processing of inference.py).

Public surface (same names as the reference where one exists):
    MOTMPNet                     models/mpn.py:144
    edge_features                inference.py:453-456
    post_processing              inference.py:70
    pruning, splitting, remove_edges_single_direction, compute_SCC_and_Clusters      utils.py
    ShardedMPN                   row-block sharded forward across the GPUs of one box (new)
    compute_P_R_F                inference.py:20-66
    evaluation.{adjusted_rand_score, adjusted_mutual_info_score, homogeneity_score, completeness_score, v_measure_score}
                                 the sklearn.metrics calls of inference.py:509-519
    tracking_table, save_mtmc    inference.py:540-551, main.py:114
"""
from . import _lib
from . import evaluation
from .evaluation import compute_P_R_F, clustering_scores
from .tracking_output import relabel_detections, save_mtmc, tracking_table
from .edge_features import edge_features
from .graph import TrackletGraph, graph_for
from .mpn import MOTMPNet
from .sharded import CudaPhases, ShardedMPN, partition_rows, shard_edges, sharded_forward
from .postprocess import (compute_SCC_and_Clusters, post_processing, pruning, remove_edges_single_direction,
                          splitting)

__all__ = ["MOTMPNet", "edge_features", "post_processing", "pruning", "splitting", "remove_edges_single_direction",
           "compute_SCC_and_Clusters", "TrackletGraph", "graph_for", "ShardedMPN", "CudaPhases", "sharded_forward", "partition_rows",
           "shard_edges", "_lib", "evaluation", "compute_P_R_F", "clustering_scores", "relabel_detections", "tracking_table",
           "save_mtmc"]
