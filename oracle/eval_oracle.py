"""TEST INFRASTRUCTURE ONLY (never imported by the product path): CPU restatement of what the reference does with the decisions
after the hot path — SURVEY.md section 8f rows 2-3.

  compute_P_R_F          inference.py:20-66, statement by statement on CPU tensors (the reference's ``.cuda()`` zero constants
                         become CPU zeros)
  relabel_loop           inference.py:540-548: the per-tracklet mask loop over the detections table
  clustering scores      the reference calls scikit-learn (0.24.2, env_gnn.yml:107; third-party, not under /root/reference):
                         the oracle is the installed scikit-learn itself
  savetxt                main.py:114 is numpy's own np.savetxt(..., fmt='%d')
Pinned by tests/golden/eval_prf.npz (outputs of the reference's compute_P_R_F run in the build container).
"""
import numpy as np
import torch


def compute_P_R_F(preds: torch.Tensor, labels: torch.Tensor):
    index_label_1 = torch.where(labels == 1)[0]                           # inference.py:21-22
    index_label_0 = torch.where(labels == 0)[0]
    precision_class1, precision_class0 = [], []
    sum_successes_1 = torch.sum(preds[index_label_1] == labels[index_label_1])       # :26-30
    if sum_successes_1 == 0:
        precision_class1.append(torch.tensor(0.0))
    else:
        precision_class1.append((sum_successes_1 / len(labels[index_label_1])) * 100.0)
    sum_successes_0 = torch.sum(preds[index_label_0] == labels[index_label_0])       # :33-37
    if sum_successes_0 == 0:
        precision_class0.append(torch.tensor(0.0))
    else:
        precision_class0.append((sum_successes_0 / len(labels[index_label_0])) * 100.0)
    TP = torch.sum(preds[index_label_1] == 1)                                         # :42-48
    FP = torch.sum(preds[index_label_0] == 1)
    TN = torch.sum(preds[index_label_0] == 0)
    FN = torch.sum(preds[index_label_1] == 0)
    P = TP / (TP + FP) if (TP + FP) != 0 else torch.tensor(0.0)                       # :50-63
    R = TP / (TP + FN) if (TP + FN) != 0 else torch.tensor(0.0)
    F = 2 * (P * R) / (P + R) if (P + R) != 0 else torch.tensor(0.0)
    return TP, FP, TN, FN, P, R, F, precision_class0, precision_class1


def relabel_loop(det_cam, det_id, node_cam, node_old, node_new):
    """inference.py:540-548 on numpy columns: masks on the ORIGINAL ids, tracklets in index order."""
    det_cam, det_id = np.asarray(det_cam), np.asarray(det_id)
    out = det_id.copy()
    for n in range(len(node_new)):
        out[(det_id == node_old[n]) & (det_cam == node_cam[n])] = int(node_new[n])
    return out


def clustering_scores(labels_true, labels_pred):
    from sklearn import metrics
    return {"adjusted_rand_score": metrics.adjusted_rand_score(labels_true, labels_pred),
            "adjusted_mutual_info_score": metrics.adjusted_mutual_info_score(labels_true, labels_pred),
            "homogeneity_score": metrics.homogeneity_score(labels_true, labels_pred),
            "completeness_score": metrics.completeness_score(labels_true, labels_pred),
            "v_measure_score": metrics.v_measure_score(labels_true, labels_pred)}
