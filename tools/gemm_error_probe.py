#!/usr/bin/env python
"""Error of the GEMM paths against float64 (max, rms and SIGNED mean relative to |ref|): 0 production tcgen05
(split accumulator for K >= 1024), 3 split accumulator forced, 4 single accumulator forced, 1 fp32 SIMT."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch

import vrpx

dev = vrpx.require_device()
for (R, K, NOUT) in ((8192, 1024, 128), (8192, 512, 128), (8192, 128, 384), (8192, 128, 512)):
    g = torch.Generator().manual_seed(K + NOUT)
    X = (torch.randn(R, K, generator=g) * 0.7 + 0.2).to(dev)
    W = (torch.randn(NOUT, K, generator=g) / K ** 0.5).to(dev)
    ref = X.double() @ W.double().T
    for path in (1, 4, 3, 0):
        Y = torch.empty(R, NOUT, device=dev)
        vrpx.check(vrpx.lib().vrpx_debug_gemm(vrpx.ptr(X), R, K, vrpx.ptr(W), NOUT, None, 0, None, None, None, vrpx.ptr(Y), path,
                                              vrpx.stream_ptr(dev)))
        torch.cuda.synchronize()
        d = Y.double() - ref
        big = ref.abs() > 0.5 * ref.abs().mean()
        rel = (d / ref)[big]
        print(f"R={R} K={K} NOUT={NOUT} path={path}: max|d|/scale {d.abs().max().item() / ref.abs().max().item():.2e} "
              f"rms rel {rel.pow(2).mean().sqrt().item():.2e} signed mean rel {rel.mean().item():+.2e} "
              f"mean(d*sign(ref))/mean|ref| {((d * ref.sign()).mean() / ref.abs().mean()).item():+.2e}", flush=True)
