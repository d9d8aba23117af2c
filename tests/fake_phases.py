"""CPU stand-in for the per-rank compute of the sharded forward (tests only).

Implements the same phase interface and the same 96-double moment-sum layout as libmpn_b200's plan API, in torch
fp64, following the hoisted algebra of DESIGN.md.  Used to exercise the N>1 schedule (row-block partitioning, moment
all-reduces, h all-gather) with gloo on CPU, and as an independent check that the hoisted/closed-form formulation
equals the reference formulation (oracle.mpn_forward).
"""
import torch

from oracle import mpn_oracle as mo

ENC0, ENC1, EDGE, NODE, APPLY = range(5)
EPS = mo.BN_EPS


class FakePhases:
    def __init__(self, sd, params, x, edge_index_local, edge_attr_local, n0, n1, total_edges, n_out):
        f = torch.float64
        self.sd = {k: v.to(f) for k, v in sd.items()}
        self.lay = mo.model_layouts(params, "resnet101")
        self.x = x.to(f)
        self.row, self.col = edge_index_local[0].long(), edge_index_local[1].long()
        self.ea = edge_attr_local.to(f)
        self.n0, self.n1, self.n = n0, n1, float(total_edges)
        self._sums = torch.zeros(96, dtype=f)
        self._h = torch.zeros(x.shape[0], 32, dtype=f)
        self.logits = torch.zeros(max(n_out, 1), self.row.numel(), 2, dtype=f)
        e = "encoder.edge_mlp.fc_layers"
        self.W1, self.b1, self.g1, self.be1 = (self.sd[f"{e}.0.weight"], self.sd[f"{e}.0.bias"], self.sd[f"{e}.1.weight"], self.sd[f"{e}.1.bias"])
        self.W2, self.b2, self.g2, self.be2 = (self.sd[f"{e}.4.weight"], self.sd[f"{e}.4.bias"], self.sd[f"{e}.5.weight"], self.sd[f"{e}.5.bias"])
        m = "MPNet.edge_model.edge_mlp.fc_layers"
        self.We, self.bE, self.g3, self.be3 = (self.sd[f"{m}.0.weight"], self.sd[f"{m}.0.bias"], self.sd[f"{m}.1.weight"], self.sd[f"{m}.1.bias"])
        m = "MPNet.node_model.node_mlp.fc_layers"
        self.Wn, self.bN, self.g4, self.be4 = (self.sd[f"{m}.0.weight"], self.sd[f"{m}.0.bias"], self.sd[f"{m}.1.weight"], self.sd[f"{m}.1.bias"])
        c = "classifier.edge_mlp.fc_layers"
        self.Wc, self.bc = self.sd[f"{c}.0.weight"], self.sd[f"{c}.0.bias"]
        self.y = None

    # -- interface ---------------------------------------------------------------------------------
    def sums(self):
        return self._sums

    def h_full(self):
        return self._h

    def node_encoder(self):
        self._h.copy_(mo.mlp_forward(self.sd, "encoder.node_mlp.fc_layers", self.lay["encoder.node_mlp.fc_layers"], self.x))

    def reduce(self, stage, with_consts=False):
        pass                                    # sums are produced fully reduced by sweep()

    def _a1(self):
        return torch.relu(self.ea @ self.W1f.t() + self.c1)

    def _e0(self):
        return torch.relu(self._a1() @ self.W2f.t() + self.c2)

    def sweep(self, step, stage, out_index=None, last=False):
        s = self._sums
        s.zero_()
        if stage == ENC0:
            a, b = self.ea[:, 0], self.ea[:, 1]
            s[0], s[1], s[2], s[3], s[4] = a.sum(), b.sum(), (a * a).sum(), (a * b).sum(), (b * b).sum()
        elif stage == ENC1:
            u = self._a1() @ self.W2.t() + self.b2
            s[0:4], s[4:8] = u.sum(0), (u * u).sum(0)
        elif stage == EDGE:
            ein = self._e0() if step == 1 else torch.relu(self.y * self.s3 + self.t3)
            self.y = self.Ps[self.row - self.n0] + self.Pd[self.col] + ein @ self.We[:, 64:68].t()
            s[0:4], s[4:8] = self.y.sum(0), (self.y * self.y).sum(0)
        elif stage == NODE:
            ep = torch.relu(self.y * self.s3 + self.t3)
            lr = self.row - self.n0
            nloc = self.n1 - self.n0
            S1 = torch.zeros(nloc, 4, dtype=ep.dtype).index_add_(0, lr, ep)
            deg = torch.bincount(lr, minlength=nloc).to(ep.dtype)
            w = self.Wn[:, 32:36]                                  # [32,4]
            q = S1 @ w.t()                                         # [nloc,32]
            s[0:32] = (deg[:, None] * self.A + q).sum(0)
            s[32:64] = (deg[:, None] * self.A * self.A + 2 * self.A * q).sum(0)
            T2 = ep.t() @ ep
            idx = 0
            for a in range(4):
                for b in range(a, 4):
                    s[64 + idx] = T2[a, b]
                    idx += 1
        elif stage == APPLY:
            if step == 0:                                          # L == 0: classify the encoder output
                self.logits[out_index] = self._e0() @ self.Wc.t() + self.bc
                return
            ep = torch.relu(self.y * self.s3 + self.t3)
            lr = self.row - self.n0
            z = (self.A * self.s4 + self.t4)[lr] + ep @ (self.Wn[:, 32:36] * self.s4[:, None]).t()
            self.msg = torch.zeros(self.n1 - self.n0, 32, dtype=ep.dtype).index_add_(0, lr, torch.relu(z))
            if out_index is not None:
                self.logits[out_index] = ep @ self.Wc.t() + self.bc

    def finalize(self, step, stage):
        s, n = self._sums, self.n
        if stage == ENC0:
            ma, mb = s[0] / n, s[1] / n
            caa, cab, cbb = s[2] / n - ma * ma, s[3] / n - ma * mb, s[4] / n - mb * mb
            w0, w1 = self.W1[:, 0], self.W1[:, 1]
            mean = w0 * ma + w1 * mb + self.b1
            var = w0 * w0 * caa + 2 * w0 * w1 * cab + w1 * w1 * cbb
            sc = self.g1 / torch.sqrt(var + EPS)
            self.W1f, self.c1 = self.W1 * sc[:, None], sc * self.b1 + self.be1 - sc * mean
        elif stage in (ENC1, EDGE):
            mean = s[0:4] / n
            var = s[4:8] / n - mean * mean
            if stage == ENC1:
                sc = self.g2 / torch.sqrt(var + EPS)
                self.W2f, self.c2 = self.W2 * sc[:, None], sc * self.b2 + self.be2 - sc * mean
            else:
                self.s3 = self.g3 / torch.sqrt(var + EPS)
                self.t3 = self.be3 - self.s3 * mean
        elif stage == NODE:
            w = self.Wn[:, 32:36]
            T2 = torch.zeros(4, 4, dtype=s.dtype)
            idx = 0
            for a in range(4):
                for b in range(a, 4):
                    T2[a, b] = T2[b, a] = s[64 + idx]
                    idx += 1
            quad = ((w @ T2) * w).sum(1)
            mean = s[0:32] / n
            var = (s[32:64] + quad) / n - mean * mean
            self.s4 = self.g4 / torch.sqrt(var + EPS)
            self.t4 = self.be4 - self.s4 * mean

    def node_tables(self, step):
        h = self._h
        self.Ps = h[self.n0:self.n1] @ self.We[:, 0:32].t() + self.bE
        self.Pd = h @ self.We[:, 32:64].t()
        self.A = h[self.n0:self.n1] @ self.Wn[:, 0:32].t() + self.bN

    def node_finalize(self, step):
        self._h[self.n0:self.n1] = self.msg


class FakePostOps:
    """CPU stand-in for ``CudaPostOps`` (shard compaction / rounds on the merged active list / write-back), backed by the
    post-processing oracle, so the exchange logic of ``sharded_post_processing`` can run under gloo.  A fake "graph" is any
    object with ``src`` / ``dst`` (global int64 ids of the shard's edges, edge order)."""

    def compact(self, graph, pred, prob1):
        idx = torch.nonzero(pred != 0).reshape(-1)
        return (graph.src[idx].to(torch.int32), graph.dst[idx].to(torch.int32), idx.to(torch.int32), prob1[idx].float())

    def run(self, num_cameras, src, dst, prob, CONFIG, n_nodes, numbering):
        import numpy as np
        from oracle import postproc_oracle as po
        labels, act = po.post_processing_rounds(src.numpy().astype(np.int64), dst.numpy().astype(np.int64),
                                                np.ones(src.numel(), dtype=np.int64), prob.numpy(), num_cameras, n_nodes,
                                                cutting=CONFIG['CUTTING'], pruning=CONFIG['PRUNING'],
                                                splitting=CONFIG['SPLITTING'], numbering=numbering)
        return torch.from_numpy(np.asarray(labels).astype(np.int64)), torch.from_numpy(act.astype(np.uint8))

    def clear(self, pred, eid, keep):
        pred[eid.long()[keep == 0]] = 0
