"""profiles/r2_scaling_final.md from the bench lines of 1 / 2 / 4 / 8 GPUs:
    python tools/scaling_table.py profiles/r2_bench_final_n1.json profiles/r2_bench_final_2gpu.json ... > profiles/r2_scaling_final.md
Efficiency here is value(N) / (N x value(1)) of these files (the driver computes its own from its own runs)."""
import json
import sys

lines = []
for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().split("\n")[-1])
    lines.append((d["n_gpus"], path, d))
lines.sort(key=lambda t: t[0])
base = next((d for n, _, d in lines if n == 1), None)
print("# Round 2, final: weak scaling of `bench.py` (per-GPU edges fixed at 14.7 M) and strong scaling of BASELINE configs[4]\n")
print("Each row is one `bench.py --gpus N` run (torchrun, one rank per GPU, fused peer-memory path, shared symmetric Gram); times are device")
print("times (CUDA events, max over ranks), 200 timed steps, L2 flushed between steps, sharded decisions checked against the unsharded run.\n")
print("| GPUs | tracklets | directed edges | ms/step | G edges/s | efficiency vs 1 GPU | e2e ms/graph (host in, decisions out) | decisions differing outside the margin band | file |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for n, path, d in lines:
    eff = d["value"] / (n * base["value"]) if base else float("nan")
    wl = d["config"]["workload"]
    tracklets = wl.split(" tracklets")[0].split()[-1]
    edges = wl.split("E=")[1].split()[0]
    par = d.get("parity") or {}
    print("| %d | %s | %s | %.4f | %.1f | %.2f | %.3f | %s | `%s` |" % (n, tracklets, edges, d["ms_per_step"], d["value"] / 1e9, eff, d["e2e"]["ms_per_step"],
                                                                   par.get("decisions_differ_outside_margin_band"), path.split("/")[-1]))
print("\nTimeline of one step per N (`step_timeline_ms`: CUDA events at the phase boundaries of the forward call, ms since it began, max over ranks):\n")
keys = ["edge_features+enc0_end", "node_encoder_end(side stream)", "enc1_end", "joined", "h_arrived", "node_tables_end", "edge_update_end",
        "node_moments_end", "node_apply_end", "node_finalize_end"]
print("| GPUs | " + " | ".join(k.replace("_end", "") for k in keys) + " | step - forward (graph build + host) |")
print("|---:|" + "---:|" * (len(keys) + 1))
for n, _, d in lines:
    tl = d.get("step_timeline_ms") or {}
    row = ["%.3f" % tl[k] if k in tl else "—" for k in keys]
    rest = d["ms_per_step"] - tl.get("node_finalize_end", float("nan"))
    print("| %d | " % n + " | ".join(row) + " | %.3f |" % rest)
print("\nStrong scaling of BASELINE configs[4] (N = 32,768 tracklets, 8 cameras, L = 4, 939,524,096 directed edges, tables from camera ids):\n")
print("| GPUs | ms/step | G edges/s | speed-up vs 1 GPU | efficiency | GB per GPU | decisions differing outside the margin band vs unsharded |")
print("|---:|---:|---:|---:|---:|---:|---:|")
one = None
for n, _, d in lines:
    c5 = d.get("c5_strong") if n > 1 else (d.get("extra") or {}).get("configs4_big_graph_1gpu")
    if not c5 or "ms_per_step" not in c5:
        continue
    if n == 1:
        one = c5["ms_per_step"]
    sp = one / c5["ms_per_step"] if one else float("nan")
    print("| %d | %.2f | %.1f | %.2f | %.2f | %.1f | %s |" % (n, c5["ms_per_step"], c5["edges_per_s"] / 1e9, sp, sp / n, c5.get("peak_mem_gb_rank0", float("nan")),
                                                        (c5.get("vs_unsharded") or {}).get("decisions_differ_outside_margin_band", "—")))
