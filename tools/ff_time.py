#!/usr/bin/env python
"""Time the fused feed-forward kernel (csrc/ff_fused.cu) at the C4 row count, with and without the converters' arithmetic,
and the two-GEMM path it replaces."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch

import vrpx

dev = vrpx.require_device()
L = vrpx.lib()
R = 65536 * 50
X = torch.randn(R, 128, device=dev)
W1 = torch.randn(512, 128, device=dev) / 128 ** 0.5
b1 = torch.randn(512, device=dev) * 0.1
W2 = torch.randn(128, 512, device=dev) / 512 ** 0.5
b2 = torch.randn(128, device=dev) * 0.1
sc = torch.rand(128, device=dev) + 0.5
sh = torch.randn(128, device=dev)
Y = torch.empty_like(X)
H = torch.empty(R, 512, device=dev)
st = vrpx.stream_ptr(dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fused():
    vrpx.check(L.vrpx_debug_ff_fused(vrpx.ptr(X), R, vrpx.ptr(W1), vrpx.ptr(b1), vrpx.ptr(W2), vrpx.ptr(b2), vrpx.ptr(X), vrpx.ptr(sc),
                                     vrpx.ptr(sh), vrpx.ptr(Y), st))


def two_gemms():
    vrpx.check(L.vrpx_debug_gemm(vrpx.ptr(X), R, 128, vrpx.ptr(W1), 512, vrpx.ptr(b1), 1, None, None, None, vrpx.ptr(H), 0, st))
    vrpx.check(L.vrpx_debug_gemm(vrpx.ptr(H), R, 512, vrpx.ptr(W2), 128, vrpx.ptr(b2), 0, vrpx.ptr(X), vrpx.ptr(sc), vrpx.ptr(sh), vrpx.ptr(Y), 0, st))


flops = 2.0 * R * 128 * 512 * 2
for dbg in (0, 4, 5):
    L.vrpx_debug_encoder_fuse_ff(1 | (dbg << 1))
    ms = timed(fused)
    print(f"ff_fused dbg={dbg}: {ms:.3f} ms  useful {flops / ms / 1e9:.0f} TFLOP/s ({3 * flops / ms / 1e9:.0f} issued)", flush=True)
L.vrpx_debug_encoder_fuse_ff(1)
ms = timed(two_gemms)
print(f"FF1 + FF2 GEMMs: {ms:.3f} ms", flush=True)
