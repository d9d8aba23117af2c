#!/usr/bin/env python
"""Benchmark of the MPN hot path (BASELINE.json metric: MPN inference directed edges/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step = one pass of the hot path over one synthetic graph that is already resident in HBM:
graph tables (K0) -> edge features (K1) -> MOTMPNet.forward with fused decisions (K1b-K4).
N = 1: BASELINE.json configs[1] — L=1 MPN, 4096 tracklets, 8 cameras, dense cross-camera edges, E = 14,680,064
directed edges ("~7M" undirected pairs).  N > 1: the same per-GPU work (weak scaling): a graph of 4096*sqrt(N) tracklets
whose edges are sharded by row block, one all-reduce of BatchNorm moment sums per BatchNorm.
Prints ONE JSON line (rank 0).  `--impl reference` times the reference algorithm's CPU restatement (oracle/) on the
host cores instead (the reference itself is pure Python on ATen and cannot travel to the GPU box).
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints a "NCCL version ..." banner through C stdio when
# the first communicator is created), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to a
# private duplicate of the original stdout.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()

import torch  # noqa: E402

NODES_1GPU, CAMS, FEAT_DIM = 4096, 8, 2048
METRIC, UNIT = "mpn_inference_directed_edges_per_sec", "edges/s"
# algorithmic work per directed edge (SURVEY.md section 8d / DESIGN.md "Kernels")
# node_apply: y read 16 + logits 8 + the fused decisions this bench asks for (u8 prediction 1 + fp32 probability 4), SURVEY 8d row K4
BYTES_PER_EDGE = {"enc_moments": 16.0, "edge_update": 28.0, "node_moments": 16.0, "node_apply": 29.0, "forward": 89.0}
FLOP_PER_EDGE_GRAM = 4096.0


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: an NVML polling thread (2 ms period; the timed region of
    the default run is only tens of ms, too short for `nvidia-smi -lms`), with `nvidia-smi` as the fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        import threading
        self.samples, self.bits, self.max_mhz, self.power = [], 0, None, []
        self.stop_flag = threading.Event()
        self.thread = self.p = self.f = None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                i = 0
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        if i % 4 == 0:
                            self.bits |= int(reasons_fn(h))
                            self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                    except Exception:
                        pass
                    i += 1
                    time.sleep(0.001)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.how = "nvml thread, 2 ms period"
        except Exception:
            self.how = "nvidia-smi -lms 20"
            try:
                self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                           "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
                time.sleep(1.0)                                   # nvidia-smi needs ~0.5 s before its first sample
            except Exception:
                self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "how": self.how}
        reasons = set()
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            reasons = {name for name, bit in self.REASONS if self.bits & bit}
            if self.power:
                out["power_w_max"] = max(self.power)
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()
            self.f.flush()
            rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
            os.unlink(self.f.name)
            for r in rows:
                try:
                    self.samples.append(float(r[0]))
                    out["sm_max_mhz"] = float(r[1])
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                        if v.strip().lower().startswith("active"):
                            reasons.add(name)
                except Exception:
                    pass
        sm = sorted(self.samples)
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------- synthetic graph
def workload_shape(world):
    """(tracklets, directed edges) of the bench workload on ``world`` GPUs: configs[1] at 1 GPU, E per GPU ~ constant above."""
    if world == 1:
        n = NODES_1GPU
    else:
        n = int(round(NODES_1GPU * world ** 0.5 / (CAMS * world))) * CAMS * world           # weak scaling
    return n, n * (n - n // CAMS)                                                            # equal cameras, all cross-camera pairs


def workload_name(n_nodes, e_total, world):
    return ("BASELINE configs[1]: L=1 MPN (shipped config) + edge features + decisions, %d tracklets, %d cameras, dense "
            "cross-camera edges, E=%d directed edges%s" %
            (n_nodes, CAMS, e_total, "" if world == 1 else " row-block sharded over %d GPUs" % world))


def device_graph(n_nodes, cams, seed, dev, row_block=None):
    """Graph(N,C,seed) of SURVEY.md section 8d built on the device.  row_block=(n0,n1) builds only that shard's edges."""
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.nn.functional.normalize(torch.randn(n_nodes, FEAT_DIM, generator=g, device=dev), p=2, dim=0)   # inference.py:403-404
    cam = (torch.arange(n_nodes, device=dev) * cams // n_nodes)
    nodes = torch.arange(n_nodes, device=dev)
    n0, n1 = row_block if row_block is not None else (0, n_nodes)
    parts = []
    for c in range(cams):
        rows = nodes[(cam == c) & (nodes >= n0) & (nodes < n1)]
        if rows.numel():
            parts.append(torch.cartesian_prod(rows, nodes[cam != c]))         # inference.py:409-413
    ei = torch.cat(parts, dim=0).t().contiguous()
    return x, ei


def make_model(dev, L=1, n_cls=1):
    import gcn_mtmc_b200 as m
    from oracle.mpn_oracle import init_weights, shipped_model_params     # synthetic weights only (not the measured path)
    params = shipped_model_params(L, n_cls)
    net = m.MOTMPNet(copy.deepcopy(params), None, "resnet101")
    net.load_state_dict(init_weights(params, "resnet101", 0), strict=True)
    net = net.to(dev).eval()
    net.fuse_decisions = True
    return net


class Batch:
    pass


# ------------------------------------------------------------------------------------------------- CPU baseline
def cpu_reference_rate(seconds_budget=20.0, nodes=2048, cams=8, reps=1):
    """Reference algorithm on the host cores (oracle/ = plain-torch restatement of models/mpn.py + inference.py:453-456).
    Bounded sample: full forward on a (nodes, cams) graph + edge features on a 200k-edge slice, scaled per edge."""
    from oracle import mpn_oracle as mo
    torch.set_num_threads(os.cpu_count() or 1)
    params = mo.shipped_model_params(1, 1)
    sd = mo.init_weights(params, "resnet101", 0)
    x, ei, _, _ = mo.synth_graph(nodes, cams, 0)
    E = ei.shape[1]
    n_ef = min(E, 200_000)
    t0 = time.perf_counter()
    ea_part = mo.edge_features(x, ei[:, :n_ef], chunk=50_000)
    t_ef = (time.perf_counter() - t0) / n_ef
    ea = torch.empty(E, 2)
    ea[:, 0] = ea_part[:, 0].mean()
    ea[:, 1] = ea_part[:, 1].mean()
    ea[:n_ef] = ea_part
    ea += 0.01 * torch.randn(E, 2, generator=torch.Generator().manual_seed(1))
    best = None
    with torch.no_grad():
        for _ in range(max(1, reps)):
            t0 = time.perf_counter()
            mo.mpn_forward(sd, params, "resnet101", x, ei, ea)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            if time.perf_counter() - t0 > seconds_budget:
                break
    per_edge = best / E + t_ef
    return {"value": 1.0 / per_edge, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "oracle/mpn_oracle.py (torch-CPU restatement of the reference): full L=1 forward on N=%d C=%d E=%d "
                      "(%.2f s) + reference edge-feature ops on a %d-edge slice (%.2f us/edge), per-edge costs added" %
                      (nodes, cams, E, best, n_ef, t_ef * 1e6),
            "forward_edges_per_sec": E / best, "edge_feature_edges_per_sec": 1.0 / t_ef}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    info = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        info = cpu_reference_rate(nodes=1024 if i < args.warmup else 2048)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    v = info["value"]
    world = max(int(args.gpus), 1)
    n_nodes, e_total = workload_shape(world)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / max(len(times), 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n_nodes, e_total, world),
                       "sample": "each step = one bounded CPU sample of that workload: the same model and graph family at "
                                 "2048 tracklets x %d cameras, rate per directed edge (see cpu_baseline.sample)" % CAMS},
            "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------- GPU arm
def time_phases(m, net, x, ei, reps=5):
    """Per-kernel device times through the plan API (CUDA events on the launching stream), L2 flushed between reps."""
    import ctypes as C
    dev = x.device
    g = m.TrackletGraph(ei, x.shape[0])
    ea = m.edge_features(x, ei, graph=g)
    W = net._weights(dev)
    logits = torch.empty(1, g.n_edges, 2, device=dev)
    pred = torch.empty(g.n_edges, dtype=torch.uint8, device=dev)
    prob1 = torch.empty(g.n_edges, device=dev)
    ph = m.CudaPhases(g, W, x, ea, 1, 1, g.n_edges, logits, pred, prob1, True, ws_kind="bench_plan")
    S = m._lib
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    seq = [("node_encoder", lambda: ph.node_encoder()),
           ("enc_moments0", lambda: ph.sweep(0, S.STAGE_ENC0)), ("fin0", lambda: ph.reduce(S.STAGE_ENC0, True)),
           ("enc_moments1", lambda: ph.sweep(0, S.STAGE_ENC1)), ("fin1", lambda: ph.reduce(S.STAGE_ENC1, True)),
           ("node_tables", lambda: ph.node_tables(1)),
           ("edge_update", lambda: ph.sweep(1, S.STAGE_EDGE)), ("fin2", lambda: ph.reduce(S.STAGE_EDGE, True)),
           ("node_moments", lambda: ph.sweep(1, S.STAGE_NODE)), ("fin3", lambda: ph.reduce(S.STAGE_NODE, True)),
           ("node_apply", lambda: ph.sweep(1, S.STAGE_APPLY, out_index=0, last=True)),
           ("node_finalize", lambda: ph.node_finalize(1))]
    acc = {k: 0.0 for k, _ in seq}
    for rep in range(reps + 1):
        for name, fn in seq:
            flush.fill_(rep & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            if rep > 0:
                acc[name] += a.elapsed_time(b) / reps
    ph.close()
    # edge-feature pieces
    ef = {}
    for rep in range(reps + 1):
        flush.fill_(rep & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        m.edge_features(x, ei, graph=g)
        b.record()
        b.synchronize()
        if rep > 0:
            ef["edge_features"] = ef.get("edge_features", 0.0) + a.elapsed_time(b) / reps
    acc.update(ef)
    # the Gram GEMM alone (tcgen05, 3xFP16 planes, symmetric tiles: the kernel the edge features run), through the exported block
    N, D = x.shape
    L = S.lib()
    G = torch.empty(N, N, device=dev)
    gws = torch.empty(L.mpn_gemm_nt_workspace_bytes(N, N, D, 1), dtype=torch.uint8, device=dev)
    amax = x.abs().max().reshape(1).float()
    t_gemm = 0.0
    for rep in range(reps + 1):
        flush.fill_(rep & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        S.check(L.mpn_gram_nt(x.data_ptr(), 0, G.data_ptr(), N, N, D, amax.data_ptr(), gws.data_ptr(), gws.numel(),
                              torch.cuda.current_stream().cuda_stream))
        b.record()
        b.synchronize()
        if rep > 0:
            t_gemm += a.elapsed_time(b) / reps
    acc["gram_gemm"] = t_gemm
    return acc


def s02_latency(m, dev, reps=200):
    """Second half of BASELINE.json's metric: p50 latency of one S02-shaped graph (configs[0]: N=300, C=4, E=67,500, L=1):
    graph tables + edge features + forward + fused decisions, device time per call (CUDA events), host wall clock beside it."""
    net = make_model(dev)
    x, ei = device_graph(300, 4, 0, dev)
    b = Batch()
    b.x, b.edge_index, b.num_nodes = x, ei, 300

    def call():
        g = m.TrackletGraph(ei, 300, validate="deferred")
        b._mpn_b200_graph = ((ei.data_ptr(), tuple(ei.shape), ei._version, 300, None), g)
        b.edge_attr = m.edge_features(x, ei, graph=g)
        net(b)
        g.validate()
        return net.last_pred
    for _ in range(20):
        call()
    dts, hts = [], []
    for _ in range(reps):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        a.record()
        call()
        e.record()
        e.synchronize()
        hts.append(1e3 * (time.perf_counter() - t0))
        dts.append(a.elapsed_time(e))
    dts.sort()
    hts.sort()
    return {"config": "BASELINE configs[0] shape: 300 tracklets, 4 cameras, E=%d directed edges, L=1; K0 + K1 + forward + decisions per call "
                      "(forward replayed as one CUDA graph)" % ei.shape[1],
            "p50_ms": dts[len(dts) // 2], "p99_ms": dts[int(len(dts) * 0.99)], "host_wall_p50_ms": hts[len(hts) // 2], "calls": reps}


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import gcn_mtmc_b200 as m
    m._lib.require_device(local_rank)
    lib = m._lib.lib()
    peaks = load_peaks()
    net = make_model(dev)

    n_nodes, e_expected = workload_shape(world)
    per = n_nodes // world
    blocks = [(r * per, (r + 1) * per) for r in range(world)]
    x, ei = device_graph(n_nodes, CAMS, 0, dev, row_block=blocks[rank])
    E_local = ei.shape[1]
    tot = torch.tensor([E_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    E_total = int(tot.item())
    sharded = m.ShardedMPN(net) if world > 1 else None
    batch = Batch()
    batch.x, batch.edge_index, batch.num_nodes = x, ei, n_nodes

    def step(x, ei):
        if world == 1:
            g = m.TrackletGraph(ei, n_nodes, validate="deferred")              # K0 (the edge-list check is read at the end)
            batch.x, batch.edge_index = x, ei
            batch._mpn_b200_graph = ((ei.data_ptr(), tuple(ei.shape), ei._version, n_nodes, None), g)
            batch.edge_attr = m.edge_features(x, ei, graph=g)                  # K1
            out, h = net(batch)                                                # K1b..K4 (+ fused decisions)
            g.validate()                                                       # raises on an unsorted / out-of-range edge list
            return net.last_pred
        n0, n1 = blocks[rank]
        g = m.TrackletGraph(ei, n_nodes, row_offset=n0, n_rows=n1 - n0, validate="deferred")
        ea = m.edge_features(x, ei, graph=g)
        out, h, pred, prob1 = sharded.forward(x, ei, ea, blocks, fuse_decisions=True, graph=g)
        g.validate()
        return pred

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank) if rank == 0 else None       # polls through warm-up + timed steps (same kernels; an NVML
    for _ in range(max(args.warmup, 3)):                              # query takes milliseconds, the timed region tens of them)
        step(x, ei)
    barrier()
    launches0 = lib.mpn_kernel_launches()
    total_ms = 0.0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                                                  # L2 flush between timed iterations
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step(x, ei)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)                          # max over ranks
        total_ms += float(ms.item())
    launches = lib.mpn_kernel_launches() - launches0
    clk = clocks.stop() if clocks else None
    ms_per_step = total_ms / args.steps
    value = E_total / (ms_per_step * 1e-3)

    # ---- e2e: public API with HOST buffers (pinned), copies inside the timed region.
    # Primary: what the reference driver holds on the host before it builds the graph (inference.py:383-414): the node
    # features and the per-node camera ids; the graph tables are built on the device (TrackletGraph.from_cameras, row f1).
    # Secondary ("int64_edge_index"): the caller ships the reference's int64 edge_index [2,E] over PCIe as well.
    hx = x.cpu().pin_memory()
    hei = ei.cpu().pin_memory()
    hpred = torch.empty(E_local, dtype=torch.uint8).pin_memory()
    dx, dei = torch.empty_like(x), torch.empty_like(ei)
    cam_host = (torch.arange(n_nodes) * CAMS // n_nodes).numpy()
    n_e2e = max(3, min(args.steps, 11))

    def e2e_loop(fn):
        samples = []
        for i in range(n_e2e + 1):
            flush.fill_(i & 0xFF)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pred = fn()
            hpred.copy_(pred, non_blocking=True)
            b.record()
            barrier()
            ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if i > 0:
                samples.append(float(ms.item()))
        samples.sort()
        return samples[len(samples) // 2]                  # median over the timed calls (max over ranks each): robust to a host hiccup

    def e2e_edge_index():
        dx.copy_(hx, non_blocking=True)
        dei.copy_(hei, non_blocking=True)
        return step(dx, dei)

    def e2e_cameras():
        dx.copy_(hx, non_blocking=True)
        if world == 1:
            g = m.TrackletGraph.from_cameras(cam_host, dev)                # K0 from camera ids, on the device
            batch.x, batch.mpn_graph = dx, g
            batch.edge_attr = None                                         # edge features inside forward (overlapped with the encoder)
            net(batch)
            return net.last_pred
        g = m.TrackletGraph.from_cameras(cam_host, dev, row_block=blocks[rank])
        ea = m.edge_features(dx, None, graph=g)
        return sharded.forward(dx, None, ea, blocks, fuse_decisions=True, graph=g)[2]

    e2e_ei_ms = e2e_loop(e2e_edge_index)
    e2e_ms = e2e_loop(e2e_cameras)
    batch.mpn_graph = None
    # Throughput of a STREAM of host-resident graphs (the reference loops over one-graph batches, inference.py:375): GraphStream
    # keeps two graphs in flight so the PCIe copies of neighbouring graphs overlap the kernels.  Every graph still pays its own
    # H2D (features + camera ids) and D2H (decisions) inside the timed region; the region closes after the last D2H.
    pipe_ms, pipe_depth, n_pipe = None, 2, max(args.steps, 3)
    if world == 1:
        gs = m.GraphStream(net, dev, depth=pipe_depth, graph_replay=os.environ.get("MPN_BENCH_GRAPH_REPLAY") == "1")   # experimental knob
        hpreds = [torch.empty(E_local, dtype=torch.uint8).pin_memory() for _ in range(pipe_depth + 1)]

        def run_pipe(k):
            for i in range(k):
                gs.submit(hx, cam_host, hpreds[i % len(hpreds)])
            gs.drain(host_sync=False)

        run_pipe(2 * pipe_depth)                  # every slot used twice (the experimental replay mode captures on the second use)
        flush.fill_(1)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_pipe(n_pipe)
        b.record()
        barrier()
        pipe_ms = a.elapsed_time(b) / n_pipe
        gs.drain()
        if not torch.equal(hpreds[(n_pipe - 1) % len(hpreds)], hpred):
            raise RuntimeError("GraphStream decisions differ from the one-at-a-time call")
    h2d = hx.numel() * 4 + cam_host.size * 8
    d2h = hpred.numel()

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": workload_name(n_nodes, E_total, world),
                           "l2": "256 MiB flush between timed iterations; inputs (edge_index 235 MB) exceed L2",
                           "timing": "CUDA events per step on the launching stream, max over ranks, summed over steps"},
                "e2e": {"value": E_total / ((pipe_ms or e2e_ms) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "ms_per_step": pipe_ms or e2e_ms,
                        "inputs": "host node features [N,2048] f32 + camera ids (graph tables built on the device; one "
                                  "MOTMPNet.forward call with data.edge_attr=None, i.e. edge features computed inside it)",
                        "how": ("GraphStream(depth=%d): %d graphs submitted back to back from pinned host memory, timed from the first "
                                "H2D to the last D2H (CUDA events); the H2D / D2H of neighbouring graphs overlap the kernels, every "
                                "graph pays its own copies; no L2 flush (each graph's ~600 MB of edge arrays exceed L2)"
                                % (pipe_depth, n_pipe)) if pipe_ms else
                               "one graph at a time: H2D, kernels, D2H serial; median over the timed calls, max over ranks",
                        "one_at_a_time": {"value": E_total / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                                          "how": "H2D, kernels, D2H serial per call, barrier + L2 flush between calls; median"},
                        "int64_edge_index": {"value": E_total / (e2e_ei_ms * 1e-3), "ms_per_step": e2e_ei_ms,
                                             "h2d_bytes_per_step": hx.numel() * 4 + hei.numel() * 8}},
                "gpu_launches": int(launches), "clocks": clk}
        knobs = {"pdl": lib.mpn_set_pdl(-1) == 2, "fused_distance": lib.mpn_set_fused_distance(-1) == 2,
                 "apply_arrive": os.environ.get("MPN_ATC_ARRIVE", "0")[:1] == "1",
                 "graph_replay": os.environ.get("MPN_BENCH_GRAPH_REPLAY") == "1"}
        if any(knobs.values()):                        # experimental switches (off by default) label the line they produced
            line["experimental"] = knobs
    if world == 1:
        ph = time_phases(m, net, x, ei)
        fwd_ms = sum(v for k, v in ph.items() if k not in ("edge_features", "gram_gemm"))
        hbm = peaks["hbm_gbs"]
        tf32_peak = peaks["bf16_tflops"] / 2.0
        E = E_local
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if os.path.isfile(tp):
            traffic = json.load(open(tp))
        roof = {}
        for name, key in (("enc_moments", None), ("edge_update", "edge_update"), ("node_moments", "node_moments"), ("node_apply", "node_apply")):
            t = (ph["enc_moments0"] + ph["enc_moments1"]) if key is None else ph[key]
            ach = BYTES_PER_EDGE[name] * E / (t * 1e-3) / 1e9
            roof[name] = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "ms": t,
                          "traffic": traffic.get(name)}
        ach = BYTES_PER_EDGE["forward"] * E / (fwd_ms * 1e-3) / 1e9
        roof["forward_all_kernels"] = {"bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "ms": fwd_ms, "traffic": None}
        t = ph["gram_gemm"]
        ach = FLOP_PER_EDGE_GRAM * E / (t * 1e-3) / 1e12
        roof["gram_gemm"] = {"bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                             "ms": t, "traffic": traffic.get("gram_gemm"),
                             "peak_note": "peak = TF32 dense proxy = measured bf16 burst / 2 (SURVEY 8d; no TF32 measurement).  The kernel runs "
                                          "3 fp16 products per fp32 product (3xFP16, kind::f16 at the bf16 rate) on half of the tiles "
                                          "(symmetric Gram); 'achieved' counts the algorithmic 4096 FLOP per directed edge; the time "
                                          "includes the fp16 split of x (one launch before the GEMM)"}
        t = ph["edge_features"]
        roof["edge_features_all_kernels"] = {"bound": "tensor", "achieved": FLOP_PER_EDGE_GRAM * E / (t * 1e-3) / 1e12, "peak": tf32_peak,
                                             "unit": "TFLOP/s", "frac": FLOP_PER_EDGE_GRAM * E / (t * 1e-3) / 1e12 / tf32_peak, "ms": t, "traffic": None}
        single = [k for k in roof if not k.endswith("all_kernels")]
        dominant = max(single, key=lambda k: roof[k]["ms"])
        line["roofline"] = dict(roof[dominant], kernel=dominant, peak_source=peaks["source"],
                                traffic_source=traffic.get("_source"))
        line["roofline_all"] = roof
        line["phase_ms"] = ph
        line["s02_latency"] = s02_latency(m, dev)
        line["cpu_baseline"] = {k: v for k, v in cpu_reference_rate(reps=1).items()}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
