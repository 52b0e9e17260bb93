"""Times the attention backward (csrc/attention_bwd.cu, mma.sync) against the fp32 SIMT kernel at B x N."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch
import vrpx

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
dev = vrpx.require_device()
L = vrpx.lib()
qkv = torch.randn(B * N, 384, device=dev)
att = torch.randn(B * N, 128, device=dev) * 0.3
datt = torch.randn(B * N, 128, device=dev) * 1e-3
dqkv = torch.empty(B * N, 384, device=dev)
for path in (0, 1):
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vrpx.check(L.vrpx_debug_attention_backward(vrpx.ptr(qkv), vrpx.ptr(att), vrpx.ptr(datt), vrpx.ptr(dqkv), B, N, path,
                                                   vrpx.stream_ptr(dev)))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    gb = B * N * (384 + 128 + 128 + 384) * 4 / 1e9
    print(f"path {path} ({'mma.sync' if path == 0 else 'fp32 SIMT'}): {min(ts):.3f} ms, {gb:.1f} GB of operands -> {gb / min(ts) * 1e3:.0f} GB/s", flush=True)
