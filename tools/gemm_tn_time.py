"""Times the weight-gradient GEMM C += A^T B (csrc/gemm_tn_tc.cu vs the mma.sync kernel) on the encoder shapes of the
C4 train step (R = 65,536 x 50 rows)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch
import vrpx

dev = vrpx.require_device()
L = vrpx.lib()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536 * 50
for (M, N) in ((128, 512), (512, 128), (384, 128), (128, 128)):
    A = torch.randn(R, M, device=dev) * 1e-4
    Bm = torch.randn(R, N, device=dev)
    C = torch.zeros(M, N, device=dev)
    for path in (0, 1):
        L.vrpx_debug_gemm_tn_path(path)
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            vrpx.check(L.vrpx_gemm_tn_accumulate(vrpx.ptr(A), vrpx.ptr(Bm), vrpx.ptr(C), R, M, N, vrpx.stream_ptr(dev)))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        gb = R * (M + N) * 4 / 1e9
        print(f"M={M} N={N} path={path}: {ms:.3f} ms, {2 * R * M * N / ms / 1e9:.1f} TFLOP/s, operands {gb:.1f} GB -> {gb / ms * 1e3:.0f} GB/s", flush=True)
    L.vrpx_debug_gemm_tn_path(0)
    del A, Bm
