"""Accuracy / speed diagnosis of the two GEMM kernels (not a benchmark)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gcn_mtmc_b200 as m

dev = torch.device("cuda", 0)
L = m._lib.lib()


def run(A, B, impl, reps=0):
    M, K = A.shape; N = B.shape[0]
    C = torch.empty(M, N, device=dev)
    ws = torch.empty(max(L.mpn_gemm_nt_workspace_bytes(M, N, K, impl), 4096), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    m._lib.check(L.mpn_gemm_nt(A.data_ptr(), B.data_ptr(), None, C.data_ptr(), M, N, K, impl, ws.data_ptr(), ws.numel(), st))
    ms = None
    if reps:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            m._lib.check(L.mpn_gemm_nt(A.data_ptr(), B.data_ptr(), None, C.data_ptr(), M, N, K, impl, ws.data_ptr(), ws.numel(), st))
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / reps
    return C, ms


for (M, N, K) in [(512, 512, 64), (512, 512, 256), (512, 512, 1024), (512, 512, 2048), (512, 512, 8192)]:
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.randn(M, K, device=dev, generator=g); B = torch.randn(N, K, device=dev, generator=g)
    ref = A.double() @ B.double().t()
    for impl in (0, 1):
        C, _ = run(A, B, impl)
        d = C.double() - ref
        toward_zero = (d * torch.sign(ref)).mean().item()
        print("K=%5d impl=%d  max|err|=%.3e  rms=%.3e  mean(err*sign(ref))=%+.3e  max|err|/K=%.2e" %
              (K, impl, d.abs().max().item(), d.pow(2).mean().sqrt().item(), toward_zero, d.abs().max().item() / K))
for (M, N, K) in [(4096, 4096, 2048), (4096, 1024, 2048), (4096, 512, 1024), (4096, 128, 512)]:
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    for impl in (0, 1):
        _, ms = run(A, B, impl, reps=10)
        print("M=%d N=%d K=%d impl=%d  %.3f ms  %.1f TFLOP/s (algorithmic 2MNK)" % (M, N, K, impl, ms, 2.0 * M * N * K / ms / 1e9))
