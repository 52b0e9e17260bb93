"""Time the encoder GEMM shapes at the C4 row count (65,536 instances x 50 nodes) on the selectable GEMM paths
(vrpx_debug_gemm): 0 = production tcgen05, 4 = candidate, 1 = fp32 SIMT."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch
import vrpx

dev = vrpx.require_device()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536 * 50
paths = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4]
L = vrpx.lib()
shapes = [("qkv", 128, 384, False, False), ("out+res+bn", 128, 128, False, True), ("ff1+relu", 128, 512, True, False),
          ("ff2+res+bn", 512, 128, False, True), ("qk-table", 128, 768, False, False)]
for name, K, NOUT, relu, resbn in shapes:
    X = torch.randn(R, K, device=dev)
    W = torch.randn(NOUT, K, device=dev) / K ** 0.5
    b = torch.randn(NOUT, device=dev)
    res = torch.randn(R, NOUT, device=dev) if resbn else None
    sc = torch.rand(NOUT, device=dev) + 0.5 if resbn else None
    sh = torch.randn(NOUT, device=dev) if resbn else None
    Y = torch.empty(R, NOUT, device=dev)
    ref = None
    for path in paths:
        def run():
            vrpx.check(L.vrpx_debug_gemm(vrpx.ptr(X), R, K, vrpx.ptr(W), NOUT, vrpx.ptr(b), int(relu), vrpx.ptr(res),
                                         vrpx.ptr(sc), vrpx.ptr(sh), vrpx.ptr(Y), path, vrpx.stream_ptr(dev)))
        for _ in range(2):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        sample = Y[:: max(1, R // 4096)].double().clone()
        if ref is None:
            Xs = X[:: max(1, R // 4096)].double()
            r = Xs @ W.double().T + b.double()
            if relu:
                r = r.clamp_min(0)
            if resbn:
                r = (r + res[:: max(1, R // 4096)].double()) * sc.double() + sh.double()
            ref = r
        err = (sample - ref).abs().max().item()
        gb = (R * K + R * NOUT * (2 if resbn else 1)) * 4 / 1e9
        print(f"{name:12s} K={K:3d} NOUT={NOUT:3d} path {path}: {ms:7.3f} ms  {2*R*K*NOUT/ms/1e9:7.1f} TFLOP/s(fp32-equiv)  "
              f"{gb/ms*1e3:6.0f} GB/s  max err {err:.2e} (scale {ref.abs().max().item():.1f})")
