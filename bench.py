#!/usr/bin/env python
"""bench.py — throughput of the fused greedy rollout (BASELINE.json metric: instance-steps/s, greedy TSP-50).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--nodes N] [--kind tsp]

One "step" = one full greedy rollout (encoder forward + every decode step + every environment transition) over
one batch of synthetic uniform instances.  Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement.

  value      whole-job instance-steps/s with the instances already resident in HBM (Philox on device)
  e2e        the same through the public API with HOST buffers: pinned host arrays -> TSPEnv.from_arrays ->
             agent.evaluate(env) -> costs back on the host, copies inside the timed region
  roofline   the decode loop (persistent kernel for steps 0-1, then glimpse / GEMM-B / pointer launches per step):
             algorithmic bytes (512*N+100 per instance-step) / the CUDA-event time of those launches,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline / --impl reference
             the CPU restatement of the reference algorithm (oracle/, numpy env + torch-CPU policy, all host
             threads) on a bounded sample of the same workload.  The unmodified Python reference cannot travel to
             the GPU box; the restatement is pinned to it by tests/test_oracle_*.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "vrp-gym_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

KINDS = ("tsp", "vrp", "irp")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kind", default="tsp", choices=KINDS)
    ap.add_argument("--nodes", type=int, default=50)
    ap.add_argument("--batch", type=int, default=65536, help="instances per GPU (weak scaling)")
    ap.add_argument("--coupling", type=int, default=-1, help="glimpse-mask coupling group; -1 = whole per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=4096, help="instances in the bounded CPU sample")
    ap.add_argument("--gemm-path", type=int, default=0, help="0 tcgen05 f16-split (production), 1 fp32 SIMT cross-check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-split", action="store_true",
                    help="keep every decode step inside the persistent kernel (A/B against the split-step launches)")
    ap.add_argument("--seed", type=int, default=69)
    ap.add_argument("--mode", default="rollout", choices=["rollout", "train"],
                    help="rollout: greedy evaluate (headline); train: one REINFORCE step (sampled rollout + sampled "
                         "baseline rollout + backward + Adam), BASELINE.json configs[3]")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_rollout_rate(kind, N, B, seed, steps=1):
    """Greedy rollout of the oracle port on host cores.  Returns (instance-steps/s, seconds per step, threads)."""
    from agents import IRPAgent, TSPAgent, VRPAgent
    from oracle import policy_oracle as po
    from oracle.env_oracle import EnvOracle

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    rs = np.random.RandomState(seed)
    xy = rs.rand(B, N, 2)
    depot = rs.randint(0, N, size=B)
    C = 0.2449 * N + 26.12
    demand = rs.uniform(1, 10, size=(B, N)) / C
    demand[np.arange(B), depot] = 0
    agent = {"tsp": TSPAgent, "vrp": VRPAgent, "irp": IRPAgent}[kind](seed=seed)
    sd = {k: v.detach().float().cpu() for k, v in agent.model.state_dict().items()}
    total_steps, t0 = 0, time.perf_counter()
    with torch.no_grad():
        for _ in range(steps):
            env = EnvOracle(kind, xy, depot, demand)
            po.rollout(sd, env, greedy=True)
            total_steps += env.step_count * B
    dt = time.perf_counter() - t0
    return total_steps / dt, dt / steps, threads


def run_reference_arm(a):
    """--impl reference: the reference algorithm on the host CPU (oracle port), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for _ in range(min(a.warmup, 1)):
        cpu_rollout_rate(a.kind, a.nodes, min(a.cpu_batch, 64), a.seed)
    rate, sec, threads = cpu_rollout_rate(a.kind, a.nodes, a.cpu_batch, a.seed, steps=a.steps)
    sample = f"greedy {a.kind.upper()}-{a.nodes} rollout of {a.cpu_batch} instances per step (oracle port of the reference: numpy env + torch-CPU policy)"
    line = {
        "impl": "reference", "metric": "instance_steps_per_sec", "value": rate, "unit": "instance-steps/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"greedy {a.kind.upper()}-{a.nodes} rollout, CPU sample batch {a.cpu_batch}"},
        "cpu_baseline": {"value": rate, "unit": "instance-steps/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "instance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            if not (t0 <= ts <= t1 + 0.1):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ train step
def run_train_bench(a, Env, Agent, dev, rank, world, dist):
    """One step = the body of TSPAgent.train's epoch without baseline_update (graph_tsp_agent.py:176-186): reset,
    sampled rollout with grad (train-mode BatchNorm), sampled baseline rollout, REINFORCE loss, backward, Adam."""
    import vrpx

    B, N = a.batch, a.nodes
    agent = Agent(seed=a.seed)
    agent.model.encoder.gemm_path = agent.target_model.encoder.gemm_path = a.gemm_path
    env = Env(N, B, 0, seed=a.seed, instance_rng="philox", instance_offset=rank * B)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def one():
        agent.model.train()
        loss_m, loss_b, logp = agent.step(env, (False, True))
        adv = (loss_m - loss_b) * -1
        loss = agent.policy_gradient_step(adv, logp)
        return env.step_count, float(loss)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        T, loss = one()
    barrier()
    l0 = vrpx.launch_count()
    e0, e1 = ev(), ev()
    w0 = time.perf_counter()
    e0.record()
    inst_steps = 0
    for _ in range(a.steps):
        T, loss = one()
        inst_steps += 2 * T * B
    e1.record()
    barrier()
    w1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(inst_steps)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms, inst_steps = float(t.item()), float(tot.item())
        line = {"metric": "train_instance_steps_per_sec", "value": inst_steps / (ms * 1e-3), "unit": "instance-steps/s",
                "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"REINFORCE train step {a.kind.upper()}-{N}: sampled rollout + sampled baseline rollout "
                                       f"+ backward + Adam, {B} instances per GPU", "instances_per_gpu": B, "nodes": N},
                "loss": loss, "clocks": clocks.stop(w0, w1), "gpu_launches": int(vrpx.launch_count() - l0),
                "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    a = parse()
    if a.impl == "reference":
        return run_reference_arm(a)

    import vrpx
    from agents import IRPAgent, TSPAgent, VRPAgent
    from agents.graph_encoder import run_encoder
    from gym_vrp.envs import IRPEnv, TSPEnv, VRPEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    Env = {"tsp": TSPEnv, "vrp": VRPEnv, "irp": IRPEnv}[a.kind]
    Agent = {"tsp": TSPAgent, "vrp": VRPAgent, "irp": IRPAgent}[a.kind]
    B, N = a.batch, a.nodes
    if a.mode == "train":
        return run_train_bench(a, Env, Agent, dev, rank, world, dist)
    agent = Agent(seed=a.seed)  # identical seed-initialised weights on every rank (no checkpoint in the reference tree)
    model = agent.model
    model.eval()
    model.encoder.gemm_path = a.gemm_path
    model.coupling = None if a.coupling < 0 else a.coupling
    env = Env(N, B, 0, seed=a.seed, instance_rng="philox", instance_offset=rank * B)  # shard = own reference batch

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def one_step(events=None):
        env.restart_episode()
        if events is not None:
            events[0].record()
        with torch.no_grad():
            h = run_encoder(model.encoder, env=env, depot=env._depot if model._USES_DEPOT_EMBED else None,
                            gemm_path=a.gemm_path)
            if events is not None:
                events[1].record()
            out = model.decoder.rollout_episode(env, h, greedy=True, coupling=model.coupling)
            if events is not None:
                events[2].record()
        return out

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        out = one_step()
    T = out["steps"]
    barrier()
    launches0 = vrpx.launch_count()
    per_step_events = [[ev(), ev(), ev()] for _ in range(a.steps)]
    e0, e1 = ev(), ev()
    w0 = time.perf_counter()
    e0.record()
    total_inst_steps = 0
    kernel_ms = []
    if a.no_split:
        vrpx.lib().vrpx_debug_rollout_split(0)
    vrpx.lib().vrpx_debug_rollout_timing(1)  # CUDA events around the decode launches alone, on their launch stream
    for i in range(a.steps):
        out = one_step(per_step_events[i])
        total_inst_steps += out["steps"] * B  # includes the .item() sync on the step count
        kernel_ms.append(float(vrpx.lib().vrpx_debug_rollout_kernel_ms()))  # the step is already synchronised
    e1.record()
    barrier()
    w1 = time.perf_counter()
    elapsed_ms = e0.elapsed_time(e1)
    launches = vrpx.launch_count() - launches0
    enc_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in per_step_events]))
    roll_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in per_step_events]))
    mean_cost = float(out["cost"].mean().item())
    vrpx.lib().vrpx_debug_rollout_timing(0)
    kern_ms = float(np.mean(kernel_ms))

    # ---- e2e: public API with host buffers (pinned H2D of the instances, D2H of the costs inside the timed region)
    s = env.sampler
    xy_h = torch.from_numpy(s.get_graph_positions()).pin_memory()
    dep_h = torch.from_numpy(s.get_depots()[:, 0].astype(np.int64)).pin_memory()
    dem_h = torch.from_numpy(s.get_demands()[:, :, 0]).pin_memory()

    def e2e_step():
        e = Env.from_arrays(xy_h.numpy(), dep_h.numpy(), dem_h.numpy() if a.kind != "tsp" else None, device=dev)
        loss = agent.evaluate(e)
        return loss.cpu(), e.step_count

    for _ in range(2):
        e2e_step()
    barrier()
    f0, f1 = ev(), ev()
    f0.record()
    e2e_inst_steps = 0
    n_e2e = max(1, min(a.steps, 3))
    for _ in range(n_e2e):
        loss_h, sc = e2e_step()
        e2e_inst_steps += sc * B
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    h2d = int(xy_h.numel() * 8 + dep_h.numel() * 4 + (dem_h.numel() * 8 if a.kind != "tsp" else 0))
    d2h = int(B * 4)

    # ---- max over ranks
    t = torch.tensor([elapsed_ms, e2e_ms, roll_ms, enc_ms, kern_ms], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(total_inst_steps), float(e2e_inst_steps)], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    elapsed_ms, e2e_ms, roll_ms, enc_ms, kern_ms = [float(x) for x in t.tolist()]
    total_inst_steps, e2e_inst_steps = [float(x) for x in tot.tolist()]

    if rank == 0:
        clk = clocks.stop(w0, w1)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, burst)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        alg_bytes = (512.0 * N + 100.0) * B * T          # per rollout-kernel launch
        # DRAM bytes of one launch from the committed `ncu --set full` capture of this exact config, else null
        traffic = None
        tpath = os.path.join(ROOT, "profiles", f"rollout_traffic_{a.kind}{N}_b{B}.json")
        if os.path.exists(tpath):
            traffic = float(json.load(open(tpath))["dram_bytes_per_launch"])
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        value = total_inst_steps / (elapsed_ms * 1e-3)
        line = {
            "metric": "instance_steps_per_sec", "value": value, "unit": "instance-steps/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": elapsed_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (tensor-core contractions on f16 hi/lo splits with f32 accumulation, ~fp32 accuracy; env f64/bitmask)"
            if a.gemm_path == 0 else "f32", "data": "synthetic",
            "config": {"workload": f"greedy {a.kind.upper()}-{N} rollout (encoder + {T} fused decode/env steps), "
                                   f"{B} Philox-uniform instances per GPU, seed-initialised {a.kind.upper()}Agent weights",
                       "instances_per_gpu": B, "nodes": N, "steps_per_rollout": T,
                       "coupling_group": B if model.coupling is None else model.coupling,
                       "l2": "inputs larger than L2 (embeddings %.2f GB per GPU re-streamed every decode step)" % (B * N * 512 / 1e9)},
            "rollouts_per_sec": value / T,
            "mean_cost": mean_cost,
            "breakdown_ms": {"encoder": enc_ms, "score_tables": roll_ms - kern_ms, "decode_loop": kern_ms},
            "clocks": clk,
            "e2e": {"value": e2e_inst_steps / (e2e_ms * 1e-3), "unit": "instance-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": n_e2e},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": ("k_rollout (persistent decoder+env, every step)" if a.no_split else
                                                     "decode loop: k_rollout (steps 0-1) + per step k_step_glimpse, k_gemm_tc4 "
                                                     "(GEMM-B), k_step_pointer"), "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": kern_ms},
        }
        if not a.no_cpu_baseline:
            rate, sec, threads = cpu_rollout_rate(a.kind, N, a.cpu_batch, a.seed)
            line["cpu_baseline"] = {"value": rate, "unit": "instance-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"one greedy {a.kind.upper()}-{N} rollout of {a.cpu_batch} instances "
                                              f"({sec:.1f} s) with the oracle port (numpy env + torch-CPU policy)"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
