// mt19937_legacy.cu — the reference's instance stream, off the Python loop.
//
// The reference draws every instance from numpy's LEGACY GLOBAL RandomState (MT19937), one graph after the other:
//   coordinates  np.random.rand(N, 2)                          gym_vrp/graph/vrp_graph.py:29
//   depots       np.random.choice(N, size=D, replace=False)    vrp_graph.py:34     (= permutation(N)[:D])
//   demand       np.random.uniform(1, 10, size=(N, 1)) / C     vrp_graph.py:41-43  (C = 0.2449 N + 26.12, depot -> 0)
// after ONE np.random.choice(B, num_draw, replace=False) in the env constructor (gym_vrp/envs/tsp.py:48,55), and
// RandomAgent draws np.random.choice(feasible, 1) per instance and step (agents/random_agent.py:33-35).
// The stream is sequential by construction (the shuffle uses rejection sampling, so the number of words an instance
// consumes is data dependent): it runs on the HOST, in C, on a copy of the generator state that the Python side takes
// from / returns to numpy (np.random.get_state / set_state), so numpy and this code continue each other's stream.
//
// numpy is a third-party dependency of the reference (environment.yml:9, unpinned; the legacy stream is frozen by NumPy
// policy, NEP 19).  Restated algorithms (numpy/random/src/mt19937/mt19937.c, src/distributions/distributions.c,
// mtrand.pyx), each checked bit-for-bit against numpy 2.3 by tests/test_abi_and_host.py:
//   next_uint32     MT19937 with the standard tempering; state = key[624] + pos
//   next_double     (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53                               (random_standard_uniform)
//   interval(max)   smallest mask >= max, redraw (next_uint32 & mask) until <= max           (random_interval)
//   shuffle         for i = n-1 .. 1: j = interval(i); swap(x[i], x[j])                      (RandomState._shuffle_raw)
//   uniform(lo, r)  lo + r * next_double                                                     (random_uniform)
//   choice(pop, 1)  (replace=True) = randint(0, len): masked rejection on rng = len - 1, no draw when rng == 0
//                   (random_bounded_uint64_fill, use_masked = legacy)
// Compiled with -ffp-contract=off: `1 + 9 u` and `0.2449 N + 26.12` must round twice like numpy's.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397;

struct Mt {
  uint32_t* key;
  int pos;
  void refill() {
    const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
    int i;
    uint32_t y;
    for (i = 0; i < MT_N - MT_M; ++i) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + MT_M] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    for (; i < MT_N - 1; ++i) {
      y = (key[i] & UPPER) | (key[i + 1] & LOWER);
      key[i] = key[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    }
    y = (key[MT_N - 1] & UPPER) | (key[0] & LOWER);
    key[MT_N - 1] = key[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
    pos = 0;
  }
  inline uint32_t next_u32() {
    if (pos == MT_N) refill();
    uint32_t y = key[pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  inline double next_double() {
    const int32_t a = (int32_t)(next_u32() >> 5), b = (int32_t)(next_u32() >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
  inline uint64_t interval(uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    mask |= mask >> 32;
    uint64_t v;
    if (max <= 0xffffffffull) {
      while ((v = (next_u32() & mask)) > max) {
      }
    } else {
      while ((v = ((((uint64_t)next_u32()) << 32 | next_u32()) & mask)) > max) {
      }
    }
    return v;
  }
};

inline bool state_ok(const uint32_t* key, const int32_t* pos) { return key && pos && *pos >= 0 && *pos <= MT_N; }

}  // namespace

extern "C" {

int vrpx_mt19937_seed(uint32_t seed, uint32_t* key, int32_t* pos) {
  VRPX_CHECK_ARG(key && pos, "NULL state");
  for (int i = 0; i < MT_N; ++i) {   // init_genrand (np.random.seed(int) -> mt19937_seed)
    key[i] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
  }
  *pos = MT_N;
  return VRPX_OK;
}

int vrpx_mt19937_permutation_head(uint32_t* key, int32_t* pos, int64_t n, int64_t k, int64_t* h_out) {
  VRPX_CHECK_ARG(state_ok(key, pos) && n >= 1 && k >= 0 && k <= n && (k == 0 || h_out), "bad argument");
  Mt g{key, *pos};
  std::vector<int64_t> x((size_t)n);
  for (int64_t i = 0; i < n; ++i) x[(size_t)i] = i;
  for (int64_t i = n - 1; i >= 1; --i) {
    const int64_t j = (int64_t)g.interval((uint64_t)i);
    const int64_t t = x[(size_t)j];
    x[(size_t)j] = x[(size_t)i];
    x[(size_t)i] = t;
  }
  for (int64_t i = 0; i < k; ++i) h_out[i] = x[(size_t)i];
  *pos = g.pos;
  return VRPX_OK;
}

int vrpx_mt19937_instances(uint32_t* key, int32_t* pos, int64_t num_graphs, int32_t N, int32_t num_depots, double* h_xy,
                           int64_t* h_depots, double* h_demand) {
  VRPX_CHECK_ARG(state_ok(key, pos) && num_graphs >= 0 && N >= 1 && num_depots >= 0 && num_depots <= N, "bad argument");
  VRPX_CHECK_ARG(h_xy && h_depots && h_demand, "NULL output");
  Mt g{key, *pos};
  const double C = 0.2449 * (double)N + 26.12;   // vrp_graph.py:41
  std::vector<int64_t> perm((size_t)N);
  for (int64_t gi = 0; gi < num_graphs; ++gi) {
    double* xy = h_xy + gi * (int64_t)N * 2;
    for (int i = 0; i < 2 * N; ++i) xy[i] = g.next_double();               // vrp_graph.py:29
    for (int i = 0; i < N; ++i) perm[(size_t)i] = i;                        // :34 permutation(N)[:D]
    for (int i = N - 1; i >= 1; --i) {
      const int j = (int)g.interval((uint64_t)i);
      const int64_t t = perm[(size_t)j];
      perm[(size_t)j] = perm[(size_t)i];
      perm[(size_t)i] = t;
    }
    int64_t* dep = h_depots + gi * (int64_t)num_depots;
    for (int d = 0; d < num_depots; ++d) dep[d] = perm[(size_t)d];
    double* dem = h_demand + gi * (int64_t)N;
    for (int i = 0; i < N; ++i) {                                            // :42 uniform(1, 10) / C
      const double u = 1.0 + 9.0 * g.next_double();
      dem[i] = u / C;
    }
    for (int d = 0; d < num_depots; ++d) dem[dep[d]] = 0.0;                  // :43
  }
  *pos = g.pos;
  return VRPX_OK;
}

int vrpx_mt19937_random_actions(uint32_t* key, int32_t* pos, const double* h_mask, int64_t B, int32_t N, int64_t* h_actions) {
  VRPX_CHECK_ARG(state_ok(key, pos) && h_mask && h_actions && B >= 0 && N >= 1, "bad argument");
  Mt g{key, *pos};
  std::vector<int> feas((size_t)N);
  for (int64_t b = 0; b < B; ++b) {
    const double* m = h_mask + b * (int64_t)N;
    int c = 0;
    for (int n = 0; n < N; ++n)
      if (m[n] == 0.0) feas[(size_t)c++] = n;   // np.argwhere(mask == 0), random_agent.py:33-34
    if (c == 0) {
      vrpx::set_error("vrpx_mt19937_random_actions: instance %lld has no feasible node", (long long)b);
      *pos = g.pos;
      return VRPX_ERR_ARG;
    }
    h_actions[b] = feas[(size_t)g.interval((uint64_t)(c - 1))];   // np.random.choice(feasible, 1), :35
  }
  *pos = g.pos;
  return VRPX_OK;
}

}  // extern "C"
