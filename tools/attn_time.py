"""A/B timing of the fused QKV-projection + attention kernel (csrc/attn_fused.cu) inside the encoder at the C4 shape:
CUDA-event time of one greedy rollout minus its decode launches = encoder + score tables."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "vrp-gym_b200"), ROOT]
import torch
import vrpx
from agents import TSPAgent
from gym_vrp.envs import TSPEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
env = TSPEnv(N, B, 0, seed=69, instance_rng="philox")
agent = TSPAgent(seed=69)
L = vrpx.lib()
for fuse in (0, 1, 0, 1, 0, 1, 0, 1):
    L.vrpx_debug_encoder_fuse_attention(fuse)
    ts = []
    for _ in range(5):
        env.restart_episode()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.vrpx_debug_rollout_timing(1)
        e0.record(); agent.evaluate(env); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) - L.vrpx_debug_rollout_kernel_ms())
    print("fused attention", fuse, "encoder + tables ms", round(min(ts), 3), flush=True)
