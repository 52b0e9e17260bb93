"""One launch of the fused QKV-projection + attention kernel at B x N (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch
import vrpx

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = vrpx.require_device()
X = torch.randn(B * N, 128, device=dev)
W = torch.randn(384, 128, device=dev) / 128 ** 0.5
b = torch.randn(384, device=dev) * 0.1
att = torch.empty(B * N, 128, device=dev)
L = vrpx.lib()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vrpx.check(L.vrpx_debug_qkv_attention(vrpx.ptr(X), vrpx.ptr(W), vrpx.ptr(b), B, N, vrpx.ptr(att), vrpx.stream_ptr(dev)))
    e1.record()
    torch.cuda.synchronize()
    print("k_prepare_inproj + k_qkv_attention ms", round(e0.elapsed_time(e1), 3), flush=True)
