// env.cu — API-mode environment kernels + library-wide bookkeeping (errors, launch counter).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "env_rules.cuh"

namespace vrpx {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static int check_env(const vrpx_env* e, const char* who) {
  if (!e) {
    set_error("%s: env is NULL", who);
    return VRPX_ERR_ARG;
  }
  if (e->kind < 0 || e->kind > 2 || e->N < 2 || e->N > VRPX_MAX_NODES || e->B < 1) {
    set_error("%s: bad env header kind=%d N=%d B=%lld", who, e->kind, e->N, (long long)e->B);
    return VRPX_ERR_ARG;
  }
  if (!e->xy || !e->depot || !e->visited || !e->mask || !e->cur || !e->load) {
    set_error("%s: env has NULL arrays", who);
    return VRPX_ERR_ARG;
  }
  if (e->kind == VRPX_IRP && !e->demand) {
    set_error("%s: IRP env needs demand", who);
    return VRPX_ERR_ARG;
  }
  if (e->kind != VRPX_IRP && e->mask != e->visited) {
    set_error("%s: TSP/VRP env.mask must alias env.visited", who);
    return VRPX_ERR_ARG;
  }
  return VRPX_OK;
}

// ---------------------------------------------------------------- kernels
// One thread per (instance, node).  Distributions of vrp_graph.py:29,34,41-43 on a Philox stream.
__global__ void k_generate(vrpx_env e, uint64_t seed, uint64_t offset) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= e.B * e.N) return;
  int64_t b = idx / e.N;
  int n = (int)(idx - b * e.N);
  uint64_t gid = offset + (uint64_t)b;
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  uint4 r0 = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)n, 0u), key);
  uint4 r1 = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)n, 1u), key);
  uint4 rd = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), 0xffffffffu, 2u), key);
  int depot = (int)(((uint64_t)rd.x * (uint64_t)e.N) >> 32);
  e.xy[idx * 2 + 0] = u53(r0.x, r0.y);
  e.xy[idx * 2 + 1] = u53(r0.z, r0.w);
  if (e.demand) {
    double C = 0.2449 * (double)e.N + 26.12;
    double d = (1.0 + 9.0 * u53(r1.x, r1.y)) / C;
    e.demand[idx] = (n == depot) ? 0.0 : d;
  }
  if (n == 0) e.depot[b] = depot;
}

__global__ void k_reset(vrpx_env e) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= e.B) return;
  int depot = e.depot[b];
  Bits128 v = {{0u, 0u, 0u, 0u}};
  apply_mask_rules(v, e.kind, e.N, depot, depot);
  e.cur[b] = depot;
  e.load[b] = 1.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) e.visited[b * 4 + i] = v.w[i];
  if (e.kind == VRPX_IRP) {
    Bits128 x = demand_exceeds(e.demand + b * e.N, e.N, 1.0);
#pragma unroll
    for (int i = 0; i < 4; ++i) e.mask[b * 4 + i] = v.w[i] | x.w[i];
  }
}

__global__ void k_step(vrpx_env e, const int64_t* __restrict__ actions, double* __restrict__ reward,
                       int32_t* __restrict__ not_done) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool unfinished = false;
  if (b < e.B) {
    Bits128 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.w[i] = e.visited[b * 4 + i];
    int cur = e.cur[b];
    double load = e.load[b];
    int depot = e.depot[b];
    int a = (int)actions[b];
    a = a < 0 ? 0 : (a >= e.N ? e.N - 1 : a);  // the reference would raise IndexError; stay in bounds
    const double* dem = e.demand ? e.demand + b * e.N : nullptr;
    StepResult r = env_transition(e.kind, e.N, e.xy + b * e.N * 2, dem, depot, a, v, cur, load);
    reward[b] = -r.dist;
    unfinished = !r.all_before;
    e.cur[b] = cur;
    e.load[b] = load;
#pragma unroll
    for (int i = 0; i < 4; ++i) e.visited[b * 4 + i] = v.w[i];
    if (e.kind == VRPX_IRP) {
      Bits128 x = demand_exceeds(dem, e.N, load);
#pragma unroll
      for (int i = 0; i < 4; ++i) e.mask[b * 4 + i] = v.w[i] | x.w[i];
    }
  }
  // block-level vote, one atomic per block (global `done` = AND over the batch, tsp.py:103-104)
  int any = __syncthreads_or(unfinished ? 1 : 0);
  if (threadIdx.x == 0 && any) {
    int cnt = 0;  // exact count is not needed by callers; 1 per block is enough to signal "not done"
    cnt = 1;
    atomicAdd(not_done, cnt);
  }
}

// One thread per (instance, node): reference-layout observation (tsp.py:106-129, irp.py:101-124).
__global__ void k_observe(vrpx_env e, double* __restrict__ state, double* __restrict__ mask,
                          double* __restrict__ visited) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= e.B * e.N) return;
  int64_t b = idx / e.N;
  int n = (int)(idx - b * e.N);
  double m = (double)((e.mask[b * 4 + (n >> 5)] >> (n & 31)) & 1u);
  if (mask) mask[idx] = m;
  if (visited) visited[idx] = (double)((e.visited[b * 4 + (n >> 5)] >> (n & 31)) & 1u);
  if (state) {
    double x = e.xy[idx * 2], y = e.xy[idx * 2 + 1];
    double isd = (e.depot[b] == n) ? 1.0 : 0.0;
    if (e.kind == VRPX_IRP) {
      double* s = state + idx * 5;
      s[0] = x; s[1] = y; s[2] = e.demand[idx]; s[3] = isd; s[4] = m;
    } else {
      double* s = state + idx * 4;
      s[0] = x; s[1] = y; s[2] = isd; s[3] = m;
    }
  }
}

__global__ void k_refresh_mask(vrpx_env e) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= e.B) return;
  Bits128 x = demand_exceeds(e.demand + b * e.N, e.N, e.load[b]);
#pragma unroll
  for (int i = 0; i < 4; ++i) e.mask[b * 4 + i] = e.visited[b * 4 + i] | x.w[i];
}

__global__ void k_set_visited(vrpx_env e, const double* __restrict__ visited) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= e.B) return;
  Bits128 v = {{0u, 0u, 0u, 0u}};
  for (int n = 0; n < e.N; ++n)
    if (visited[b * e.N + n] != 0.0) bit_set(v, n);
#pragma unroll
  for (int i = 0; i < 4; ++i) e.visited[b * 4 + i] = v.w[i];
  if (e.kind == VRPX_IRP) {
    Bits128 x = demand_exceeds(e.demand + b * e.N, e.N, e.load[b]);
#pragma unroll
    for (int i = 0; i < 4; ++i) e.mask[b * 4 + i] = v.w[i] | x.w[i];
  }
}

}  // namespace vrpx

using namespace vrpx;

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

extern "C" {

int vrpx_abi_version(void) { return VRPX_ABI_VERSION; }
const char* vrpx_last_error(void) { return g_err; }
int64_t vrpx_launch_count(void) { return (int64_t)g_launches.load(); }

int vrpx_device_check(int device) {
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n <= 0) {
    set_error("vrpx_device_check: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(err));
    return VRPX_ERR_DEVICE;
  }
  if (device < 0 || device >= n) {
    set_error("vrpx_device_check: device %d out of range (%d devices)", device, n);
    return VRPX_ERR_DEVICE;
  }
  int major = 0, minor = 0;
  VRPX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  VRPX_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  if (major != 10) {
    set_error("vrpx_device_check: device %d is sm_%d%d; libvrpx is built for sm_100a only", device, major, minor);
    return VRPX_ERR_DEVICE;
  }
  return VRPX_OK;
}

int vrpx_env_generate(const vrpx_env* env, uint64_t seed, uint64_t offset, void* stream) {
  int rc = check_env(env, __func__);
  if (rc) return rc;
  VRPX_DEVICE_GUARD(env->xy);
  int64_t n = env->B * env->N;
  k_generate<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*env, seed, offset);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

int vrpx_env_reset(const vrpx_env* env, void* stream) {
  int rc = check_env(env, __func__);
  if (rc) return rc;
  VRPX_DEVICE_GUARD(env->xy);
  k_reset<<<grid_for(env->B, 256), 256, 0, (cudaStream_t)stream>>>(*env);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

int vrpx_env_observe(const vrpx_env* env, double* state, double* mask, double* visited, void* stream) {
  int rc = check_env(env, __func__);
  if (rc) return rc;
  VRPX_DEVICE_GUARD(env->xy);
  if (!state && !mask && !visited) return VRPX_OK;
  int64_t n = env->B * env->N;
  k_observe<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(*env, state, mask, visited);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

int vrpx_env_step(const vrpx_env* env, const int64_t* actions, double* reward, int32_t* not_done,
                  double* state, void* stream) {
  int rc = check_env(env, __func__);
  if (rc) return rc;
  VRPX_DEVICE_GUARD(env->xy);
  VRPX_CHECK_ARG(actions && reward && not_done, "actions/reward/not_done must be non-NULL");
  k_step<<<grid_for(env->B, 256), 256, 0, (cudaStream_t)stream>>>(*env, actions, reward, not_done);
  VRPX_LAUNCH_CHECK();
  if (state) return vrpx_env_observe(env, state, nullptr, nullptr, stream);
  return VRPX_OK;
}

int vrpx_env_refresh_mask(const vrpx_env* env, void* stream) {
  int rc = check_env(env, __func__);
  if (rc) return rc;
  if (env->kind != VRPX_IRP) return VRPX_OK;
  VRPX_DEVICE_GUARD(env->xy);
  k_refresh_mask<<<grid_for(env->B, 256), 256, 0, (cudaStream_t)stream>>>(*env);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

int vrpx_env_set_visited(const vrpx_env* env, const double* visited, void* stream) {
  int rc = check_env(env, __func__);
  if (rc) return rc;
  VRPX_DEVICE_GUARD(env->xy);
  VRPX_CHECK_ARG(visited, "visited must be non-NULL");
  k_set_visited<<<grid_for(env->B, 256), 256, 0, (cudaStream_t)stream>>>(*env, visited);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // extern "C"
