"""TSPEnv — batched travelling-salesman environment with device-resident state.

Same surface as the reference class (gym_vrp/envs/tsp.py:11-187): constructor arguments and order, `step`,
`reset`, `get_state`, `generate_mask`, `is_done`, `generate_graphs`, `render`, `enable_video_capturing`, and
the attributes `visited`, `current_location`, `depots`, `sampler`, `draw_idxs`, `step_count`.

What differs is where the state lives: instances and episode state are struct-of-arrays tensors on the GPU
(coordinates f64, depot i32, visited bitmask 4 x u32, current node i32, load f64) and every transition is
one CUDA kernel (libvrpx `vrpx_env_step`: visited update, f64 edge length, reward, mask rules, global done).
The observation is materialised in the reference layout only when a caller asks for it.  There is no CPU
implementation: constructing an environment without an sm_100 GPU raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple, Union

import numpy as np
import torch

import vrpx

from ..graph.vrp_network import VRPNetwork
from .common import ObsType


class TSPEnv:
    metadata = {"render.modes": ["human", "rgb_array"]}

    _KIND = vrpx.TSP
    _PLOT_DEMAND = False
    _STATE_COLS = 4

    def __init__(self, num_nodes: int = 20, batch_size: int = 128, num_draw: int = 6, seed: int = 69,
                 *, device=None, instance_rng: str = "numpy", instance_offset: int = 0):
        """
        Args mirror the reference (tsp.py:27-33).  Extra keyword-only arguments:
            device: CUDA device (default: current).
            instance_rng: "numpy" — instances from the legacy global numpy stream exactly like the
                reference (seed-compatible, drawn on the host); "philox" — drawn on the device
                (`vrpx_env_generate`, for large batches; not numpy-seed compatible).
            instance_offset: global id of this shard's first instance (philox mode, multi-GPU sharding).
        """
        assert num_draw <= batch_size, "Num_draw needs to be equal or lower than the number of generated graphs."
        assert 2 <= num_nodes <= vrpx.MAX_NODES, f"num_nodes must be in [2, {vrpx.MAX_NODES}]"
        assert instance_rng in ("numpy", "philox")
        self._device = vrpx.require_device(device)

        np.random.seed(seed)  # the reference seeds the *global* stream here and never again (tsp.py:48)

        self.step_count = 0
        self.num_nodes = num_nodes
        self.batch_size = batch_size
        self.seed = seed
        self._instance_rng = instance_rng
        self._instance_offset = int(instance_offset)
        self._philox_epoch = 0

        self.draw_idxs = np.random.choice(batch_size, num_draw, replace=False)
        self.video_save_path = None

        self.generate_graphs()

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_arrays(cls, xy, depots, demand=None, num_draw: int = 0, *, device=None):
        """Build an environment around given instances (host arrays): xy (B,N,2) f64, depots (B,), demand (B,N).

        Contiguous float64 arrays are ADOPTED as the host-side instance store (no host copy) and uploaded straight from
        the caller's memory; when that memory is pinned the upload is an asynchronous DMA on the current stream."""
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        B, N, _ = xy.shape
        self = cls.__new__(cls)
        self._device = vrpx.require_device(device)
        self.step_count = 0
        self.num_nodes, self.batch_size, self.seed = N, B, None
        self._instance_rng, self._instance_offset, self._philox_epoch = "numpy", 0, 0
        self.draw_idxs = np.arange(num_draw)
        self.video_save_path = None
        net = VRPNetwork(B, N, 1, plot_demand=cls._PLOT_DEMAND, _sample=False)
        net._store.xy = xy
        net._store.depots = np.asarray(depots).reshape(B, 1).astype(int)
        if demand is not None:
            net._store.demand = np.ascontiguousarray(demand, dtype=np.float64).reshape(B, N)
        self._sampler = net
        self._sampler_stale = False
        self._alloc_state()
        self._upload_instances(with_demand=demand is not None)
        self._reset_episode()
        return self

    def generate_graphs(self):
        """Draw a fresh batch of instances and reset the episode state (tsp.py:162-174)."""
        B, N = self.batch_size, self.num_nodes
        self._alloc_state()
        if self._instance_rng == "numpy":
            self._sampler = VRPNetwork(num_graphs=B, num_nodes=N, num_depots=1, plot_demand=self._PLOT_DEMAND)
            self._sampler_stale = False
            self._upload_instances()
        else:
            self._sampler = VRPNetwork(B, N, 1, plot_demand=self._PLOT_DEMAND, _sample=False)
            self._sampler_stale = True  # host copy is filled lazily from the device
            seed = np.uint64(self.seed if self.seed is not None else 0)
            off = self._instance_offset + self._philox_epoch * (1 << 40)
            vrpx.check(vrpx.lib().vrpx_env_generate(C.byref(self._view()), C.c_uint64(int(seed)),
                                                    C.c_uint64(off), vrpx.stream_ptr(self._device)))
            self._philox_epoch += 1
            self._uploaded_version = self._sampler.version
        self._reset_episode()

    def _alloc_state(self):
        B, N, dev = self.batch_size, self.num_nodes, self._device
        self._xy = torch.empty((B, N, 2), dtype=torch.float64, device=dev)
        self._depot = torch.empty((B,), dtype=torch.int32, device=dev)
        self._demand = torch.zeros((B, N), dtype=torch.float64, device=dev)
        self._visited = torch.zeros((B, 4), dtype=torch.int32, device=dev)
        self._mask = self._visited if self._KIND != vrpx.IRP else torch.zeros((B, 4), dtype=torch.int32, device=dev)
        self._cur = torch.empty((B,), dtype=torch.int32, device=dev)
        self._load = torch.ones((B,), dtype=torch.float64, device=dev)
        self._host_cur = None

    def _upload_instances(self, with_demand: bool = True):
        s = self._sampler._store
        self._xy.copy_(torch.from_numpy(s.xy), non_blocking=True)
        self._depot.copy_(torch.from_numpy(s.depots[:, 0].astype(np.int32)))
        if with_demand:
            self._demand.copy_(torch.from_numpy(s.demand), non_blocking=True)
        else:
            self._demand.zero_()
        self._uploaded_version = self._sampler.version

    def _sync_instances(self):
        """Re-upload if a caller edited instances through `sampler.graphs[i].nodes[n][...] = v`."""
        if not self._sampler_stale and self._sampler.version != self._uploaded_version:
            self._upload_instances()
            # edited demands change the IRP rule `demand - load > 0` of the CURRENT state (no-op for TSP / VRP)
            vrpx.check(vrpx.lib().vrpx_env_refresh_mask(C.byref(self._view()), vrpx.stream_ptr(self._device)))

    def _reset_episode(self):
        vrpx.check(vrpx.lib().vrpx_env_reset(C.byref(self._view()), vrpx.stream_ptr(self._device)))
        self._host_cur = None

    def _view(self) -> vrpx.EnvView:
        v = vrpx.EnvView()
        v.kind, v.N, v.B = self._KIND, self.num_nodes, self.batch_size
        v.xy, v.depot, v.demand = self._xy.data_ptr(), self._depot.data_ptr(), self._demand.data_ptr()
        v.visited, v.mask = self._visited.data_ptr(), self._mask.data_ptr()
        v.cur, v.load = self._cur.data_ptr(), self._load.data_ptr()
        return v

    # ------------------------------------------------------------------ reference attribute surface
    @property
    def sampler(self) -> VRPNetwork:
        if self._sampler_stale:  # philox instances live on the device; mirror them on first host access
            s = self._sampler._store
            s.xy[:] = self._xy.cpu().numpy()
            s.depots[:, 0] = self._depot.cpu().numpy()
            s.demand[:] = self._demand.cpu().numpy()
            self._sampler_stale = False
            self._uploaded_version = self._sampler.version
        return self._sampler

    @property
    def depots(self) -> np.ndarray:
        return self._depot.cpu().numpy().astype(int)[:, None]

    @property
    def current_location(self) -> np.ndarray:
        return self._cur.cpu().numpy().astype(int)[:, None]

    @property
    def visited(self) -> np.ndarray:
        out = torch.empty((self.batch_size, self.num_nodes), dtype=torch.float64, device=self._device)
        vrpx.check(vrpx.lib().vrpx_env_observe(C.byref(self._view()), None, None, vrpx.ptr(out),
                                               vrpx.stream_ptr(self._device)))
        return out.cpu().numpy()

    @visited.setter
    def visited(self, value):
        v = torch.as_tensor(np.ascontiguousarray(value, dtype=np.float64)).to(self._device)
        vrpx.check(vrpx.lib().vrpx_env_set_visited(C.byref(self._view()), vrpx.ptr(v), vrpx.stream_ptr(self._device)))

    # ------------------------------------------------------------------ gym-style API
    def step(self, actions: np.ndarray) -> Tuple[ObsType, np.ndarray, bool, dict]:
        """One transition for every instance (tsp.py:60-101).  actions: (batch_size, 1) node ids.
        Returns (state, reward (B,) f64 negative edge lengths, done, None)."""
        actions = np.asarray(actions)
        assert actions.shape[0] == self.batch_size, "Number of actions need to equal the number of generated graphs."
        self._sync_instances()
        self.step_count += 1
        dev = self._device
        a_host = np.ascontiguousarray(actions.reshape(self.batch_size), dtype=np.int64)
        a_dev = torch.from_numpy(a_host).to(dev, non_blocking=False)
        B, N = self.batch_size, self.num_nodes
        state = torch.empty((B, N, self._STATE_COLS), dtype=torch.float64, device=dev)
        reward = torch.empty((B,), dtype=torch.float64, device=dev)
        not_done = torch.zeros((1,), dtype=torch.int32, device=dev)
        vrpx.check(vrpx.lib().vrpx_env_step(C.byref(self._view()), vrpx.ptr(a_dev), vrpx.ptr(reward),
                                            vrpx.ptr(not_done), vrpx.ptr(state), vrpx.stream_ptr(dev)))
        if len(self.draw_idxs):  # edge bookkeeping only for the graphs that can be rendered
            prev = self._host_cur if self._host_cur is not None else None
            if prev is None:
                prev = self._cur_before_step(a_host)
            edges = np.stack([prev, a_host], axis=1)
            self._sampler.visit_edges(edges, only=[int(i) for i in self.draw_idxs])
        self._host_cur = a_host
        if self.video_save_path is not None:
            self.vid.capture_frame()
        done = np.bool_(int(not_done.item()) == 0)
        return self._wrap_state(state.cpu().numpy()), reward.cpu().numpy(), done, None

    def _cur_before_step(self, a_host):
        # first step of an episode: the vehicles start on their depots (tsp.py:173-174)
        if self.step_count == 1:
            return self._depot.cpu().numpy().astype(np.int64)
        return a_host  # unknown history (state was advanced by a fused rollout): no edge to draw

    def _wrap_state(self, state_np):
        return state_np

    def is_done(self):
        return np.all(self.visited == 1)

    def get_state(self) -> np.ndarray:
        """(batch, nodes, 4) f64: x, y, is_depot, mask (tsp.py:106-129)."""
        self._sync_instances()
        B, N = self.batch_size, self.num_nodes
        state = torch.empty((B, N, self._STATE_COLS), dtype=torch.float64, device=self._device)
        vrpx.check(vrpx.lib().vrpx_env_observe(C.byref(self._view()), vrpx.ptr(state), None, None,
                                               vrpx.stream_ptr(self._device)))
        return self._wrap_state(state.cpu().numpy())

    def generate_mask(self):
        """(batch, nodes) f64 of 0/1: 1 = node cannot be visited next (tsp.py:131-148).  The rules are applied
        on the device after every transition, so this only materialises the current mask."""
        self._sync_instances()
        out = torch.empty((self.batch_size, self.num_nodes), dtype=torch.float64, device=self._device)
        vrpx.check(vrpx.lib().vrpx_env_observe(C.byref(self._view()), None, vrpx.ptr(out), None,
                                               vrpx.stream_ptr(self._device)))
        return out.cpu().numpy()

    def reset(self) -> Union[ObsType, Tuple[ObsType, dict]]:
        """Fresh instances from the continuing stream (no reseed) and a reset episode (tsp.py:150-160)."""
        self.step_count = 0
        self.generate_graphs()
        return self.get_state()

    def restart_episode(self):
        """Reset the episode on the *same* instances (not in the reference; used by benchmarks)."""
        self.step_count = 0
        self._sync_instances()
        self._reset_episode()

    # ------------------------------------------------------------------ rendering (outside the accelerated path)
    def render(self, mode: str = "human"):
        return self.sampler.draw(self.draw_idxs)

    def enable_video_capturing(self, video_save_path: str):
        self.video_save_path = video_save_path
        if self.video_save_path is not None:
            from gym.wrappers.monitoring.video_recorder import VideoRecorder  # optional dependency

            self.vid = VideoRecorder(self, self.video_save_path)
            self.vid.frames_per_sec = 1
