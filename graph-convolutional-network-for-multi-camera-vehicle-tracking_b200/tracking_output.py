"""Decision -> tracking output: inference.py:540-551 and main.py:114 on the GPU.

The reference rewrites the ``id`` column of the detections table with one pandas boolean-mask scan per tracklet
(``data_tracking.loc[(det.id == ID_old) & (det.id_cam == CAM_ID), 'id'] = ID_new``, N scans over M rows); the masks are taken
on the ORIGINAL table, so the loop is a pure map keyed by ``(id_cam, id)`` in which the last tracklet with a given key wins.
Here it is one hash join on the device.  ``save_mtmc`` writes the text file ``np.savetxt(..., fmt='%d')`` produces.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .graph import current_stream_ptr, workspace

COLUMNS = ('id_cam', 'id', 'frame', 'xmin', 'ymin', 'width', 'height')          # inference.py:551


def _dev_i64(v, dev):
    t = v if torch.is_tensor(v) else torch.as_tensor(np.array(v))          # copy: pandas columns are read-only views
    if t.is_floating_point():
        t = t.trunc()
    return t.to(device=dev, dtype=torch.int64).reshape(-1).contiguous()


def relabel_detections(det_id_cam, det_id, cam_ids_nodes, node_labels, ID_pred, device=None):
    """New ``id`` column [M] (int64, on the device): ``ID_pred[n]`` for the last tracklet n with
    ``(cam_ids_nodes[n], node_labels[n]) == (det_id_cam[r], det_id[r])``, the old id for detections of no tracklet."""
    if device is None:
        device = ID_pred.device if torch.is_tensor(ID_pred) and ID_pred.is_cuda else torch.device("cuda", torch.cuda.current_device())
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("relabel_detections runs on a CUDA device: the B200 path has no CPU fallback")
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
    dc, di = _dev_i64(det_id_cam, dev), _dev_i64(det_id, dev)
    nc, no, nn = _dev_i64(cam_ids_nodes, dev), _dev_i64(node_labels, dev), _dev_i64(ID_pred, dev)
    if dc.numel() != di.numel():
        raise ValueError("det_id_cam and det_id must have the same length")
    if not (nc.numel() == no.numel() == nn.numel()):
        raise ValueError("cam_ids_nodes, node_labels and ID_pred must have one entry per tracklet")
    out = torch.empty_like(di)
    lib = _lib.lib()
    ws = workspace("relabel", dev, lib.mpn_relabel_workspace_bytes(nc.numel()))
    with torch.cuda.device(dev):
        _lib.check(lib.mpn_relabel_detections(dc.data_ptr(), di.data_ptr(), di.numel(), nc.data_ptr(), no.data_ptr(), nn.data_ptr(),
                                              nc.numel(), out.data_ptr(), ws.data_ptr(), ws.numel(), current_stream_ptr(dev)))
    return out


def tracking_table(data_det, cam_ids_nodes, node_labels, ID_pred, device=None):
    """``data_tracking[['id_cam','id','frame','xmin','ymin','width','height']]`` of inference.py:540-551 as an int64 array [M,7].

    ``data_det``: the detections table — a pandas DataFrame or a mapping with those seven columns."""
    cols = {k: np.asarray(data_det[k]) for k in COLUMNS}
    new_id = relabel_detections(cols['id_cam'], cols['id'], cam_ids_nodes, node_labels, ID_pred, device).cpu().numpy()
    table = np.empty((new_id.shape[0], len(COLUMNS)), dtype=np.int64)
    for j, k in enumerate(COLUMNS):
        v = new_id if k == 'id' else cols[k]
        table[:, j] = np.trunc(v).astype(np.int64) if np.issubdtype(np.asarray(v).dtype, np.floating) else v     # '%d' truncates
    return table


def save_mtmc(path, table):
    """``np.savetxt(path, table, fmt='%d')`` (main.py:114): space-separated integers, one detection per line."""
    t = np.asarray(table)
    if t.ndim != 2:
        raise ValueError("table must be 2-D")
    if np.issubdtype(t.dtype, np.floating):
        t = np.trunc(t)
    t = np.ascontiguousarray(t, dtype=np.int64)
    _lib.check(_lib.lib().mpn_write_mtmc_txt_host(str(path).encode(), t.ctypes.data, t.shape[0], t.shape[1]))
