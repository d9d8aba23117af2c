"""A/B of the launch-gap experiments on the configs[1] step, in ONE process on one GPU:

    eager                      the bench step as shipped (K0 from edge_index, edge features + forward + decisions)
    pdl                        same, programmatic dependent launch on   (needs `MPN_PDL=1 csrc/build.sh`; skipped otherwise)
    graph                      the same calls captured once as a CUDA graph and replayed
    graph+pdl                  captured with the launch attribute on (programmatic edges inside the graph)
    arrive                     eager with MPN_ATC_ARRIVE=1: the apply sweep's block barrier replaced by an mbarrier arrive
    fused[+pdl][+graph][+arrive]   the fused distance epilogue of the Gram GEMM (mpn_set_fused_distance), alone and combined
    cameras, cameras+fused...  the same with the tables built from the camera ids (bench.py's e2e path without the copies)

Every variant is checked bit for bit against the eager decisions before it is timed (L2 flushed between calls, CUDA events,
median of `reps`).  Diagnostic: prints a table and writes gpurun_out/gap_experiments.json.

    MPN_PDL=1 bash graph-convolutional-network-for-multi-camera-vehicle-tracking_b200/csrc/build.sh
    python tools/gap_experiments.py [reps [tag]]          # MPN_PDL=2 csrc/build.sh: early trigger as well (see common.cuh)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import gcn_mtmc_b200 as m


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    tag = sys.argv[2] if len(sys.argv) > 2 else ""      # suffix of the output file (e.g. the build variant)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    m._lib.require_device(0)
    lib = m._lib.lib()
    net = bench.make_model(dev)
    N = bench.NODES_1GPU
    x, ei = bench.device_graph(N, bench.CAMS, 0, dev)
    E = ei.shape[1]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(check=True):
        b = bench.Batch()
        b.x, b.edge_index, b.num_nodes, b.edge_attr = x, ei, N, None
        b.mpn_graph = g = m.TrackletGraph(ei, N, validate="deferred")
        net(b)
        if check:
            g.validate()          # as bench.py's step does (host wait, after everything is enqueued); recycles the pinned flag slot
        return net.last_pred

    cam_host = (torch.arange(N) * bench.CAMS // N).numpy()

    def step_cameras(check=True):
        """The e2e path of bench.py without its copies: tables from the camera ids (the fused distance epilogue needs them)."""
        b = bench.Batch()
        b.x, b.num_nodes, b.edge_attr = x, N, None
        b.mpn_graph = m.TrackletGraph.from_cameras(cam_host, dev)
        net(b)
        return net.last_pred

    def timeit(fn):
        ts = []
        for i in range(reps + 3):
            flush.fill_(i & 0xFF)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            e.record()
            e.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(e))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    def captured(fn=None):
        fn = fn or step
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            out = fn(check=False)                      # (the flag slot comes from the free list: no pinned allocation in the capture)
        return cg, out

    has_pdl = lib.mpn_set_pdl(-1) != 0
    lib.mpn_set_pdl(0)
    ref = step().clone()
    torch.cuda.synchronize()
    rows = {}

    def record(name, fn, out=None):
        got = fn()
        got = out if out is not None else got
        torch.cuda.synchronize()
        diff = int((got != ref).sum().item())
        l0 = lib.mpn_kernel_launches()
        med, best = timeit(fn)
        rows[name] = {"median_ms": med, "min_ms": best, "G_edges_per_s": E / med / 1e6, "decisions_that_differ": diff,
                      "launch_calls_per_step": (lib.mpn_kernel_launches() - l0) / (reps + 3)}
        print("%-30s median %.3f ms  min %.3f ms  %.2f G edges/s  decisions that differ from eager: %d" %
              (name, med, best, E / med / 1e6, diff), flush=True)

    record("eager", step)
    if has_pdl:
        lib.mpn_set_pdl(1)
        record("pdl", step)
        lib.mpn_set_pdl(0)
    else:
        print("pdl        skipped: library built without MPN_PDL=1")
    for name, on in (("graph", 0), ("graph+pdl", 1)):
        if on and not has_pdl:
            continue
        lib.mpn_set_pdl(on)
        try:
            cg, out = captured()
            record(name, cg.replay, out)
        except Exception as exc:                       # a capture that fails must not hide the other rows
            print("%-30s failed: %s" % (name, exc))
            rows[name] = {"error": str(exc)}
        finally:
            lib.mpn_set_pdl(0)
    # apply sweep: mbarrier arrive instead of the per-iteration block barrier (read from the environment at every launch)
    os.environ["MPN_ATC_ARRIVE"] = "1"
    try:
        record("arrive", step)
    except Exception as exc:
        print("%-30s failed: %s" % ("arrive", exc))
        rows["arrive"] = {"error": str(exc)}
    finally:
        os.environ["MPN_ATC_ARRIVE"] = "0"
    # the fused distance epilogue (mpn_set_fused_distance), alone and combined, on both ways of building the graph tables
    record("cameras", step_cameras)
    for base, fn in (("", step), ("cameras+", step_cameras)):
        for name, pdl, graph, arrive in (("fused", 0, 0, 0), ("fused+pdl", 1, 0, 0), ("fused+graph", 0, 1, 0),
                                         ("fused+graph+pdl", 1, 1, 0), ("fused+graph+pdl+arrive", 1, 1, 1)):
            if pdl and not has_pdl:
                continue
            name = base + name
            lib.mpn_set_fused_distance(1)
            lib.mpn_set_pdl(pdl)
            os.environ["MPN_ATC_ARRIVE"] = "1" if arrive else "0"
            try:
                if graph:
                    cg, out = captured(fn)
                    record(name, cg.replay, out)
                else:
                    record(name, fn)
            except Exception as exc:
                print("%-30s failed: %s" % (name, exc))
                rows[name] = {"error": str(exc)}
            finally:
                lib.mpn_set_fused_distance(0)
                lib.mpn_set_pdl(0)
                os.environ["MPN_ATC_ARRIVE"] = "0"
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gap_experiments%s.json" % (("_" + tag) if tag else ""), "w") as f:
        json.dump({"workload": bench.workload_name(N, E, 1), "reps": reps, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
