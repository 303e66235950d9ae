// Logging reductions and the step report as device functions of the post-physics kernel (they used
// to be most of a second, "finalize" kernel):
//
//   * termination fire counts and the number of reset envs are integer atomics, fire and forget.
//   * the episode means of the reward terms over the reset envs (reward_manager.py:205-216) are
//     accumulated as 64-bit fixed-point sums (integer atomics commute, so the logged values are
//     run-to-run deterministic) and turned into means by the last block to leave the kernel.
//   * that block also writes the step report into the host's mapped memory and, last, its sequence
//     word, on which the host spins.
//   * envs sharded over ranks: both reductions are exchanged over NVLink peer memory (plan.h PeerInbox).
//
// The ORDERED list of reset env ids (GFB_B_RESET_IDX == (terminated | truncated).nonzero(),
// managed_env.py:308-310) is not needed by the host, only by work enqueued behind this launch: the
// slabs leave their reset masks in scratch memory and compact_kernel (aux_kernels.cuh), enqueued right
// behind the post kernel, expands them while the host is still waking up from the report.
//
// RULE learnt on the way here (1M envs, config 2; the kernel without any compaction takes 110 us):
// while the memory system is saturated a global round trip costs microseconds, so NOTHING a slab does
// may wait for one before its next block barrier -- no returning atomic whose result is used at once,
// no fence behind a batch of stores.  In-kernel variants of the ordered compaction that were built and
// measured: decoupled look-back per slab right after the terminations 251 us (the resident blocks run
// in lock-step waves, every slab of a wave polled the same cache lines); two-level (per chunk of 32
// slabs, last arriver scans the chunk, look-back over chunks) at that place 156-162 us; the same with
// the chunk work moved to the end of the slab iteration 128-137 us (a finisher that waits for earlier
// chunks stalls its whole slab at the next barrier, and the late slabs convoy).
#pragma once
#include "device_utils.cuh"
#include "plan.h"

namespace gfb {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------
// order-independent sums of fp32 values: signed 64-bit fixed point with 32 fractional bits.
// Integer additions commute, so the per-slab and per-launch sums can be fire-and-forget atomics (RED,
// no return value -- a returning atomic costs the issuing warp an L2 round trip, and its slab waits
// for it at the next block barrier) and the logged means are still run-to-run deterministic.  Each
// addend is rounded to 2^-32 (2.3e-10, far below the fp32 resolution of the logged mean); the sum
// holds +-2^31.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t FX_NAN = 1u, FX_POS_INF = 2u, FX_NEG_INF = 4u;
__device__ __forceinline__ long long to_fixed(float q, uint32_t& flags) {
  if (q != q) {
    flags |= FX_NAN;
    return 0ll;
  }
  if (fabsf(q) >= 1048576.0f) {  // 2^20 per addend (x 2^11 reset envs per slab, x 2^20 slabs): treated as infinite
    flags |= q > 0.0f ? FX_POS_INF : FX_NEG_INF;
    return 0ll;
  }
  return __float2ll_rn(__fmul_rn(q, 4294967296.0f));
}
__device__ __forceinline__ double fixed_to_double(long long v, uint32_t flags) {
  if ((flags & FX_NAN) || ((flags & FX_POS_INF) && (flags & FX_NEG_INF))) return __longlong_as_double(0x7ff8000000000000ll);
  if (flags & FX_POS_INF) return __longlong_as_double(0x7ff0000000000000ll);
  if (flags & FX_NEG_INF) return __longlong_as_double((long long)0xfff0000000000000ull);
  return (double)v * 2.3283064365386963e-10;  // 2^-32
}

// ---------------------------------------------------------------------------------------------
// exchange of one small vector between the ranks (one warp; lane i owns element i < n_vals).
// Returns the sum over ranks in rank order (bit-identical on all ranks); `timed_out` is warp-uniform.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double peer_exchange(const PeerParams& pp, int kind, double mine, int n_vals, int lane,
                                                bool& timed_out) {
  const int parity = (int)(pp.seq & 1ull);
  const int me = pp.rank, W = pp.world;
  if (lane < n_vals)
    for (int p = 0; p < W; ++p) pp.inbox[p]->slot[parity][kind][me].vals[lane] = mine;
  __threadfence_system();
  __syncwarp();
  if (lane < W) st_release_sys(&pp.inbox[lane]->slot[parity][kind][me].seq, pp.seq);
  int late = 0;
  if (lane < W) {
    const unsigned long long* flag = &pp.inbox[me]->slot[parity][kind][lane].seq;
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flag) != pp.seq) {
      if (global_timer_ns() - t0 > 2000000000ull) {  // ~2 s: a peer never issued this exchange
        late = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  timed_out = __any_sync(0xffffffffu, late) != 0;
  __threadfence_system();
  double g = mine;
  if (!timed_out && lane < n_vals) {
    g = 0.0;
    for (int r = 0; r < W; ++r) g += __ldcv(&pp.inbox[me]->slot[parity][kind][r].vals[lane]);
  }
  return g;
}

// ---------------------------------------------------------------------------------------------
// the step report (one warp of the last block to leave the kernel)
//   termination_manager.py:178-182  fired fraction per term
//   managed_env.py:308-310          number of reset envs (the indices are already in GFB_B_RESET_IDX)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_report(const KParams& K, int lane) {
  const gfb_program_head& P = K.P;
  const Scratch& sc = K.s;
  const int n_t = P.n_termination, n_r = P.n_reward;
  __threadfence();
  const int n_reset = (int)ld_acquire_gpu_u32(sc.counters + CTR_TOTAL_RESET);  // every slab has added its share
  int count = 0;
  if (lane < n_t) count = atomicExch(sc.term_count + lane, 0);
  uint32_t status = 0;
  if (lane == 0) status = atomicExch(sc.status, 0u);
  double mine = lane < n_t ? (double)count : (lane == n_t ? (double)n_reset : 0.0);
  double global = mine;
  double* log_acc = GFB_BUF(double, GFB_B_LOG_ACC);
  float* log_out = GFB_BUF(float, GFB_B_LOG_OUT);
  if (log_acc && lane <= n_t) log_acc[n_r + lane] = mine;  // local partials (NCCL fallback path)
  long long denom = P.num_envs;
  if (K.peer.world > 1 && log_acc) {
    bool timed_out;
    global = peer_exchange(K.peer, PEER_KIND_COUNTS, mine, n_t + 1, lane, timed_out);
    if (timed_out) {
      if (lane == 0) status |= GFB_STATUS_PEER_TIMEOUT;
    } else {
      denom = K.peer.global_num_envs;
    }
  }
  gfb_report* rep = sc.report_host;
  if (lane < n_t) {
    rep->termination_count[lane] = count;
    rep->global_termination_count[lane] = (long long)global;
    if (log_out) log_out[n_r + lane] = fdiv((float)global, (float)denom);
  }
  if (lane == n_t) {
    rep->global_n_reset = (long long)global;
    *sc.global_reset = global;
  }
  if (lane == 0) {
    rep->n_reset = n_reset;
    rep->status = status;
  }
}
// the report is complete (and so is everything else the launch writes): tell the host
__device__ __forceinline__ void publish_report(const KParams& K, int lane) {
  __threadfence_system();
  __syncwarp();
  if (lane == 0) st_release_sys(reinterpret_cast<unsigned long long*>(&K.s.report_host->seq), K.s.report_seq);
}

// ---------------------------------------------------------------------------------------------
// logged episode means of the reward terms (one warp of the last block to leave the kernel)
//   reward_manager.py:205-216: mean over the reset envs of (episode sum / episode seconds)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_reward_means(const KParams& K, int lane) {
  const gfb_program_head& P = K.P;
  const Scratch& sc = K.s;
  const int n_r = P.n_reward;
  __threadfence();
  double sum = 0.0;
  if (lane < n_r) {
    const long long acc = (long long)atomicExch(sc.rew_acc + lane, 0ull);
    const uint32_t fl = atomicExch(sc.rew_flags + lane, 0u);
    sum = fixed_to_double(acc, fl);
  }
  double* log_acc = GFB_BUF(double, GFB_B_LOG_ACC);
  float* log_out = GFB_BUF(float, GFB_B_LOG_OUT);
  if (log_acc && lane < n_r) log_acc[lane] = sum;
  double g_reset = (double)ld_acquire_gpu_u32(sc.counters + CTR_TOTAL_RESET);
  if (K.peer.world > 1 && log_acc && n_r > 0) {
    bool timed_out;
    const double g = peer_exchange(K.peer, PEER_KIND_SUMS, sum, n_r, lane, timed_out);
    if (timed_out) {
      if (lane == 0) atomicOr(sc.status, GFB_STATUS_PEER_TIMEOUT);  // reported with the next step
    } else {
      sum = g;
      g_reset = __ldcg(sc.global_reset);
    }
  }
  if (log_out && lane < n_r) {
    const bool logged = P.reward[lane].weight != 0.0f;  // reward_manager.py:208-209
    log_out[lane] = (g_reset > 0.0 && logged) ? (float)(sum / g_reset) : 0.0f;
  }
}

}  // namespace gfb
