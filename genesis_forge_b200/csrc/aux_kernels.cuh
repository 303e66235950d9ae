// Pre-physics action kernel, finalize (ordered compaction + logging reductions), the sparse
// re-observation kernel and the stand-alone contact scatter.
#pragma once
#include "device_utils.cuh"
#include "plan.h"

namespace gfb {

// ---------------------------------------------------------------------------------------------
// action_kernel: GenesisEnv.step bookkeeping + action manager (pre-physics)
//   genesis_env.py:196       episode_length += 1
//   genesis_env.py:202-203   last_actions <- actions ; actions <- raw
//   position_action_manager.py:402-414   NaN/Inf flags; t = a*scale + offset; clamp(lo, hi)
//   position_within_limits.py:125-126    clamp(a,-1,1) * scale + offset
//   rewards.py:267-271       action_rate = sum((last_actions - actions)^2)   (consumed post-physics)
// One slab of TILE envs per block: the raw and previous action slabs are contiguous -> TMA in;
// last_actions / actions are pure copies -> TMA out of the very same shared-memory slabs.
// ---------------------------------------------------------------------------------------------
struct ActionParams {
  // the action part of the term table only (a launch copies its parameter block: 4 KB -> 0.6 KB)
  int32_t num_envs, num_dofs, action_mode, _pad;
  float action_scale[GFB_MAX_DOFS], action_offset[GFB_MAX_DOFS];
  float action_clip_lo[GFB_MAX_DOFS], action_clip_hi[GFB_MAX_DOFS];
  const float* raw_env;
  const float* raw_mgr;  // == raw_env unless a delay FIFO is active
  float* env_actions;
  float* env_last_actions;
  float* targets;
  float* action_rate;
  int32_t* episode_length;
  uint32_t* status;
  int32_t tma_ok;
  int32_t check_finite;
};

template <int TILE>
__global__ void __launch_bounds__(TILE) action_kernel(const __grid_constant__ ActionParams A) {
  extern __shared__ __align__(128) float S[];
  __shared__ __align__(8) uint64_t bar;
  const ActionParams& P = A;
  const int tid = threadIdx.x;
  const int N = P.num_envs, D = P.num_dofs;
  const int e0 = blockIdx.x * TILE;
  const int valid = min(TILE, N - e0);
  const bool active = tid < valid;
  const int e = e0 + tid;
  const bool use_tma = A.tma_ok && valid == TILE;
  const bool two_raw = A.raw_mgr != A.raw_env;

  float* s_raw = S;                   // (TILE, D) raw env actions
  float* s_prev = S + TILE * D;       // (TILE, D) previous env.actions
  float* s_tgt = S + 2 * TILE * D;    // (TILE, D) targets out
  float* s_rawm = S + 3 * TILE * D;   // (TILE, D) delayed raw (only with a FIFO)
  const uint32_t slab_bytes = (uint32_t)TILE * D * 4u;
  const size_t goff = (size_t)e0 * D;

  if (use_tma) {
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&bar, slab_bytes * (two_raw ? 3u : 2u));
      bulk_load(s_raw, A.raw_env + goff, slab_bytes, &bar);
      bulk_load(s_prev, A.env_actions + goff, slab_bytes, &bar);
      if (two_raw) bulk_load(s_rawm, A.raw_mgr + goff, slab_bytes, &bar);
    }
  } else {
    const int words = valid * D;
    for (int w = tid; w < words; w += TILE) {
      s_raw[w] = A.raw_env[goff + w];
      s_prev[w] = A.env_actions[goff + w];
      if (two_raw) s_rawm[w] = A.raw_mgr[goff + w];
    }
  }
  int ep_len = 0;
  if (active && A.episode_length) ep_len = A.episode_length[e];  // in flight while the slabs arrive
  if (use_tma && tid < 32) mbar_wait(&bar, 0);  // one warp polls, the block barrier releases the rest
  __syncthreads();
  if (active && A.episode_length) A.episode_length[e] = ep_len + 1;  // genesis_env.py:195

  // pure copies first: last_actions <- previous actions, actions <- raw
  if (use_tma) {
    if (tid == 0) {
      bulk_store(A.env_last_actions + goff, s_prev, slab_bytes);
      bulk_store(A.env_actions + goff, s_raw, slab_bytes);
      bulk_commit();
    }
  } else {
    const int words = valid * D;
    for (int w = tid; w < words; w += TILE) {
      A.env_last_actions[goff + w] = s_prev[w];
      A.env_actions[goff + w] = s_raw[w];
    }
  }

  if (active) {
    const float* a = s_raw + tid * D;
    const float* p = s_prev + tid * D;
    const float* am = (two_raw ? s_rawm : s_raw) + tid * D;
    float* t = s_tgt + tid * D;
    float rate = 0.0f;
    uint32_t status = 0;
    auto one = [&](float x, float prev, float y, int d) -> float {
      rate = add(rate, sq(sub(prev, x)));
      if (A.check_finite) {
        if (y != y) status |= GFB_STATUS_NAN_ACTION;
        if (fabsf(y) == __int_as_float(0x7f800000)) status |= GFB_STATUS_INF_ACTION;
      }
      if (P.action_mode == 2) {
        // torch.clamp_ propagates NaN
        y = (y != y) ? y : fminf(fmaxf(y, -1.0f), 1.0f);
        y = add(mul(y, P.action_scale[d]), P.action_offset[d]);
      } else {
        y = add(mul(y, P.action_scale[d]), P.action_offset[d]);
        y = (y != y) ? y : fminf(fmaxf(y, P.action_clip_lo[d]), P.action_clip_hi[d]);
      }
      return y;
    };
    if ((D & 3) == 0) {
      // rows of D floats at a 4D-byte stride: 128-bit shared accesses are conflict-free and a
      // quarter of the instructions of the scalar walk (same element order: d ascending)
      for (int d = 0; d < D; d += 4) {
        const float4 x = *reinterpret_cast<const float4*>(a + d);
        const float4 pv = *reinterpret_cast<const float4*>(p + d);
        const float4 m = two_raw ? *reinterpret_cast<const float4*>(am + d) : x;
        float4 y;
        y.x = one(x.x, pv.x, m.x, d);
        y.y = one(x.y, pv.y, m.y, d + 1);
        y.z = one(x.z, pv.z, m.z, d + 2);
        y.w = one(x.w, pv.w, m.w, d + 3);
        *reinterpret_cast<float4*>(t + d) = y;
      }
    } else {
      for (int d = 0; d < D; ++d) t[d] = one(a[d], p[d], am[d], d);
    }
    if (A.action_rate) A.action_rate[e] = rate;
    if (status) atomicOr(A.status, status);
  }

  if (A.targets && P.action_mode != 0) {
    if (use_tma) {
      fence_async_smem();
      __syncthreads();
      if (tid == 0) {
        bulk_store(A.targets + goff, s_tgt, slab_bytes);
        bulk_commit();
      }
    } else {
      __syncthreads();
      const int words = valid * D;
      for (int w = tid; w < words; w += TILE) A.targets[goff + w] = s_tgt[w];
    }
  }
  if (use_tma && tid == 0) bulk_wait_all();
}

// ---------------------------------------------------------------------------------------------
// finalize_kernel: one block.  Turns the per-slab partials of post_kernel into
//   * the ascending int64 reset index list, identical to (terminated|truncated).nonzero()
//     (managed_env.py:308-310)
//   * per-termination fire counts / fractions (termination_manager.py:178-182)
//   * per-reward-term episode means over the reset envs (reward_manager.py:205-216)
//   * the report block (n_reset, status bits)
// Reductions are fixed-order (lane-strided partial sums in double, then a shuffle tree), so the
// logged values are run-to-run deterministic.
// ---------------------------------------------------------------------------------------------
// Logging exchange between the ranks of one NVLink domain (gfb_peer_connect): every rank owns an
// inbox with one slot per sender and step parity; senders store their partials straight into the
// peers' inboxes and publish them with a release store of the exchange sequence number.
constexpr int PEER_VALS = GFB_MAX_REWARD_TERMS + GFB_MAX_TERMINATION_TERMS + 1;
struct PeerSlot {
  double vals[PEER_VALS];
  unsigned long long seq;
  unsigned long long _pad[64 - PEER_VALS - 1];
};
static_assert(sizeof(PeerSlot) == 512, "PeerSlot is padded to 512 bytes");
struct PeerInbox {
  PeerSlot slot[2][GFB_MAX_PEERS];
};
struct PeerParams {
  PeerInbox* inbox[GFB_MAX_PEERS];  // [r] = rank r's inbox (own one for r == rank)
  int32_t rank, world;              // world <= 1: single rank, no exchange
  unsigned long long seq;           // sequence number of this exchange (same on every rank)
  int64_t global_num_envs;
  uint32_t* done_counter;           // blocks of this launch that have written their results
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

struct FinalizeParams {
  Scratch s;
  PeerParams peer;
  int64_t* reset_idx;
  float* log_out;
  double* log_acc;
  int32_t tile;
  int32_t num_envs;
  int32_t n_reward;
  int32_t n_termination;
  uint32_t phases;
  uint32_t reward_weight_mask;  // bit r set: term r has weight != 0 (mean is logged)
  gfb_report* report_host;      // device address of the host-mapped report (or null)
};

constexpr int FIN_THREADS = 256;
constexpr int FIN_CHUNK_BLOCKS = 128;  // blocks that share the ordered compaction

// deterministic block-wide sum (fixed strided order per thread, shuffle tree, warps in order)
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_warp[warp] = v;
  __syncthreads();
  T total = 0;
  for (int w = 0; w < FIN_THREADS / 32; ++w) total += s_warp[w];
  return total;
}

// Grid layout: blocks [0, n_chunks) each own a contiguous range of slabs of the ordered compaction;
// blocks [n_chunks, n_chunks + n_termination) reduce one termination counter each; the following
// n_reward blocks reduce one reward term each; the last block writes n_reset / status.
__global__ void __launch_bounds__(FIN_THREADS) finalize_kernel(const FinalizeParams F, int n_chunks) {
  __shared__ int s_iwarp[FIN_THREADS / 32];
  __shared__ double s_dwarp[FIN_THREADS / 32];
  __shared__ int s_scan[FIN_THREADS];
  const int tid = threadIdx.x;
  const int nt = F.s.n_tiles;
  const int words = F.tile / 32;
  const int b = blockIdx.x;
  gfb_report* rep = F.s.report;

  if (b < n_chunks) {
    // ---- ordered compaction of this block's slab range -------------------------------------------
    const int per_block = (nt + n_chunks - 1) / n_chunks;
    const int c0 = min(b * per_block, nt), c1 = min(c0 + per_block, nt);
    // resets in all slabs before this range (every block re-reads the short count array: L2 hits)
    int before = 0;
    for (int t = tid; t < c0; t += FIN_THREADS) before += F.s.tile_reset_count[t];
    before = block_sum<int>(before, s_iwarp);
    if (!F.reset_idx) return;
    // scan the range in sweeps of FIN_THREADS slabs
    int base = before;
    for (int t0 = c0; t0 < c1; t0 += FIN_THREADS) {
      const int t = t0 + tid;
      const int cnt = t < c1 ? F.s.tile_reset_count[t] : 0;
      s_scan[tid] = cnt;
      __syncthreads();
      // Hillis-Steele inclusive scan over 256 entries
      for (int o = 1; o < FIN_THREADS; o <<= 1) {
        const int v = tid >= o ? s_scan[tid - o] : 0;
        __syncthreads();
        s_scan[tid] += v;
        __syncthreads();
      }
      int offset = base + s_scan[tid] - cnt;
      if (cnt > 0) {
        for (int w = 0; w < words; ++w) {
          uint32_t bits = F.s.tile_reset_bits[(size_t)t * words + w];
          while (bits) {
            const int bit = __ffs(bits) - 1;
            bits &= bits - 1;
            F.reset_idx[offset++] = (int64_t)t * F.tile + w * 32 + bit;
          }
        }
      }
      base += s_scan[FIN_THREADS - 1];
      __syncthreads();
    }
    return;
  }

  // total reset count (needed for the means and the report)
  int n_reset = 0;
  for (int t = tid; t < nt; t += FIN_THREADS) n_reset += F.s.tile_reset_count[t];
  n_reset = block_sum<int>(n_reset, s_iwarp);

  const int k = b - n_chunks;
  if (k < F.n_termination) {
    int acc = 0;
    if (F.phases & GFB_PHASE_TERMINATION)
      for (int t = tid; t < nt; t += FIN_THREADS) acc += F.s.tile_term_count[(size_t)k * nt + t];
    acc = block_sum<int>(acc, s_iwarp);
    if (tid == 0 && (F.phases & GFB_PHASE_TERMINATION)) {  // split execution: later launches keep the counts
      rep->termination_count[k] = acc;
      if (F.log_out) F.log_out[F.n_reward + k] = fdiv((float)acc, (float)F.num_envs);
      if (F.log_acc) F.log_acc[F.n_reward + k] = (double)acc;
    }
  } else if (k < F.n_termination + F.n_reward) {
    const int r = k - F.n_termination;
    double acc = 0.0;
    if (F.phases & GFB_PHASE_RESET)
      for (int t = tid; t < nt; t += FIN_THREADS) acc += F.s.tile_rew_sum[(size_t)r * nt + t];
    acc = block_sum<double>(acc, s_dwarp);
    if (tid == 0 && (F.phases & GFB_PHASE_RESET)) {
      const bool logged = (F.reward_weight_mask >> r) & 1u;
      const float mean = (n_reset > 0 && logged) ? (float)(acc / (double)n_reset) : 0.0f;
      rep->reward_episode_mean[r] = mean;
      if (F.log_out) F.log_out[r] = mean;
      if (F.log_acc) F.log_acc[r] = acc;
    }
  } else if (tid == 0 && (F.phases & GFB_PHASE_RESET)) {
    rep->n_reset = n_reset;
    rep->status = atomicExch(F.s.status, 0u);
    if (F.log_acc) F.log_acc[F.n_reward + F.n_termination] = (double)n_reset;
  }
  if (!(F.phases & GFB_PHASE_RESET)) return;

  // ---- global view: the LAST of the result blocks publishes counts over all ranks ----------------
  __shared__ int s_last;
  __shared__ double s_global[PEER_VALS];
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const uint32_t participants = (uint32_t)(F.n_termination + F.n_reward + 1);
    s_last = (atomicAdd(F.peer.done_counter, 1u) == participants - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) *F.peer.done_counter = 0u;
  __threadfence();
  const int n_vals = F.n_reward + F.n_termination + 1;
  bool local_only = F.peer.world <= 1 || !F.log_acc;  // (block-uniform from here on)
  if (!local_only) {
    // exchange over peer memory: my partials into every rank's inbox, then wait for everyone's
    const int parity = (int)(F.peer.seq & 1ull);
    const int me = F.peer.rank, W = F.peer.world;
    if (tid < n_vals) {
      const double mine = __ldcg(F.log_acc + tid);
      for (int p = 0; p < W; ++p) F.peer.inbox[p]->slot[parity][me].vals[tid] = mine;
    }
    __threadfence_system();
    __syncthreads();
    if (tid < W) st_release_sys(&F.peer.inbox[tid]->slot[parity][me].seq, F.peer.seq);
    int timed_out = 0;
    if (tid < W) {
      const unsigned long long* flag = &F.peer.inbox[me]->slot[parity][tid].seq;
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(flag) != F.peer.seq) {
        if (global_timer_ns() - t0 > 2000000000ull) {  // ~2 s: a peer never issued this exchange
          timed_out = 1;
          break;
        }
        __nanosleep(200);
      }
    }
    timed_out = __syncthreads_or(timed_out);
    __threadfence_system();
    if (timed_out) {
      if (tid == 0) atomicOr(&rep->status, GFB_STATUS_PEER_TIMEOUT);
      local_only = true;
    } else {
      if (tid < n_vals) {
        double g = 0.0;
        for (int r = 0; r < W; ++r) g += __ldcv(&F.peer.inbox[me]->slot[parity][r].vals[tid]);  // rank order
        s_global[tid] = g;
      }
      __syncthreads();
      const double g_reset = s_global[n_vals - 1];
      if (tid == 0) rep->global_n_reset = (int64_t)g_reset;
      if (tid < F.n_reward) {
        const bool logged = (F.reward_weight_mask >> tid) & 1u;
        if (F.log_out) F.log_out[tid] = (g_reset > 0.0 && logged) ? (float)(s_global[tid] / g_reset) : 0.0f;
      } else if (tid < F.n_reward + F.n_termination) {
        const int k = tid - F.n_reward;
        rep->global_termination_count[k] = (int64_t)s_global[tid];
        if (F.log_out) F.log_out[tid] = fdiv((float)s_global[tid], (float)F.peer.global_num_envs);
      }
    }
  }
  if (local_only) {
    if (tid == 0) rep->global_n_reset = rep->n_reset;
    if (tid < F.n_termination) rep->global_termination_count[tid] = rep->termination_count[tid];
  }
  // the finished report goes straight into the host's (mapped, pinned) copy: the host only has to
  // wait for this kernel, no separate device-to-host copy is enqueued
  if (F.report_host) {
    __threadfence();
    __syncthreads();
    const uint32_t* src = reinterpret_cast<const uint32_t*>(rep);
    uint32_t* dst = reinterpret_cast<uint32_t*>(F.report_host);
    for (int w = tid; w < (int)(sizeof(gfb_report) / 4); w += FIN_THREADS) dst[w] = __ldcg(src + w);
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------------------------
// observe_kernel: frame 0 of every observation group for a LIST of envs (or all envs), from the
// current engine state and the CACHED inverse base quaternion.  The reference observes after
// reset (managed_env.py:322-326) with post-reset engine getters but the pre-reset cached
// quaternion (entity_manager.py:134-146 vs :189-195; EntityManager.reset does not refresh it).
// ---------------------------------------------------------------------------------------------
struct ObserveHead {  // the observation part of the term table (instead of the whole 4 KB head)
  int32_t n_contact, n_obs_groups, rng_mode, _pad;
  uint64_t rng_seed, step_index;
  int32_t contact_links[GFB_MAX_CONTACT_MANAGERS];
  gfb_obs_group obs_group[GFB_MAX_OBS_GROUPS];
};

struct ObserveParams {
  ObserveHead P;
  gfb_buffers b;
  Plan plan;
  const DevObsCol* cols;
  const int64_t* idx;
  int32_t n;
};

constexpr int OBS_ENVS = 16;      // envs per block
constexpr int OBS_THREADS = 128;  // 8 threads per env for the column gather

__global__ void __launch_bounds__(OBS_THREADS) observe_kernel(const __grid_constant__ ObserveParams K) {
  extern __shared__ __align__(128) float S[];  // (OBS_ENVS, stash_stride) stash, then the column table
  __shared__ long long s_env[OBS_ENVS];
  const ObserveHead& P = K.P;
  const Plan& plan = K.plan;
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * OBS_ENVS;
  const int valid = min(OBS_ENVS, K.n - i0);
  DevObsCol* s_cols = reinterpret_cast<DevObsCol*>(S + OBS_ENVS * plan.stash_stride + 4 - ((OBS_ENVS * plan.stash_stride) & 3));
  {
    const int32_t* src = reinterpret_cast<const int32_t*>(K.cols);
    int32_t* dst = reinterpret_cast<int32_t*>(s_cols);
    const int words = plan.n_cols_total * (int)(sizeof(DevObsCol) / 4);
    for (int w = tid; w < words; w += OBS_THREADS) dst[w] = src[w];
  }
  if (tid < valid) {
    const long long e = K.idx ? (long long)K.idx[i0 + tid] : (long long)(i0 + tid);
    s_env[tid] = e;
    float* st = S + tid * plan.stash_stride;
    const float4 q = GFB_BUF(const float4, GFB_B_INV_BASE_QUAT)[e];
    const V3 iq = {q.y, q.z, q.w};
    if (plan.needs & NEED_ANG) {
      const float* v = GFB_BUF(const float, GFB_B_ANG) + e * 3;
      const V3 r = rotate(V3{v[0], v[1], v[2]}, q.x, iq);
      st[0] = r.x; st[1] = r.y; st[2] = r.z;
    }
    if (plan.needs & NEED_LIN) {
      const float* v = GFB_BUF(const float, GFB_B_VEL) + e * 3;
      const V3 r = rotate(V3{v[0], v[1], v[2]}, q.x, iq);
      st[3] = r.x; st[4] = r.y; st[5] = r.z;
    }
    if (plan.needs & NEED_GRAV) {
      const V3 r = rotate(V3{0.f, 0.f, -1.f}, q.x, iq);
      st[6] = r.x; st[7] = r.y; st[8] = r.z;
    }
    for (int m = 0; m < P.n_contact; ++m) {
      const float* cg = GFB_BUF(const float, GFB_B_CONTACTS0 + m);
      if (!cg) continue;
      cg += e * P.contact_links[m] * 3;
      for (int t = 0; t < P.contact_links[m]; ++t)
        st[plan.st_cnorm[m] + t] = norm3(cg[t * 3], cg[t * 3 + 1], cg[t * 3 + 2]);
    }
  }
  __syncthreads();
  const Philox rng(P.rng_seed);
  for (int g = 0; g < P.n_obs_groups; ++g) {
    const gfb_obs_group& og = P.obs_group[g];
    const int O = og.n_cols, OH = og.n_cols * og.history;
    const DevObsCol* cols = s_cols + og.col_begin;
    float* out = GFB_BUF(float, GFB_B_OBS_OUT0 + g);
    const float* noise = GFB_BUF(const float, GFB_B_OBS_NOISE0 + g);
    const int total = valid * O;
    for (int f = tid; f < total; f += OBS_THREADS) {
      const int row = f / O, col = f - row * O;
      const long long e = s_env[row];
      const DevObsCol d = cols[col];
      float v = 0.0f;
      if (d.kind == 1 || d.kind == 2)
        v = reinterpret_cast<const float*>(K.b.buf[d.gbuf])[e * d.row_words + d.col];
      else if (d.kind == 3)
        v = S[row * plan.stash_stride + d.a];
      v = mul(v, d.scale);
      if (d.noise != 0.f) {
        float u;
        if (P.rng_mode == 0) {
          u = noise ? noise[e * O + col] : 0.f;
        } else {
          const uint4 x = rng((uint32_t)e, (uint32_t)P.step_index, (uint32_t)(P.step_index >> 32),
                              0x1000u + (uint32_t)(og.col_begin + (col & ~3)));
          const int j = col & 3;
          const uint32_t xj = j == 0 ? x.x : (j == 1 ? x.y : (j == 2 ? x.z : x.w));
          u = sub(mul(u01(xj), 2.f), 1.f);
        }
        v = add(v, mul(u, d.noise));
      }
      out[e * OH + col] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// contact_kernel: the reference's Taichi kernel (contact/kernel.py:5-90) on its own, with its
// argument list.  One thread per env, targets outer / contact slots inner, ordered sums.
// ---------------------------------------------------------------------------------------------
__global__ void contact_kernel(const float* __restrict__ force, const float* __restrict__ position,
                               const int32_t* __restrict__ link_a, const int32_t* __restrict__ link_b,
                               const float4* __restrict__ links_quat, const int32_t* __restrict__ targets,
                               const int32_t* __restrict__ withs, float* __restrict__ out_f,
                               float* __restrict__ out_p, float* __restrict__ counts, int n_envs, int C, int L,
                               int Lc, int Lw, int has_filter) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  const int32_t* la = link_a + (size_t)e * C;
  const int32_t* lb = link_b + (size_t)e * C;
  const float* cf = force + (size_t)e * C * 3;
  const float* cp = position + (size_t)e * C * 3;
  for (int t = 0; t < Lc; ++t) {
    const int target = targets[t];
    const float4 tq = links_quat[(size_t)e * L + target];
    float fx = 0.f, fy = 0.f, fz = 0.f, px = 0.f, py = 0.f, pz = 0.f, cnt = 0.f;
    for (int c = 0; c < C; ++c) {
      const int a = la[c], b = lb[c];
      const bool is_a = a == target, is_b = b == target;
      bool hit = is_a | is_b;
      if (hit && has_filter) {
        bool keep = false;
        for (int w = 0; w < Lw; ++w) keep |= (is_a && b == withs[w]) || (is_b && a == withs[w]);
        hit = keep;
      }
      if (hit) {
        const float x = cf[c * 3], y = cf[c * 3 + 1], z = cf[c * 3 + 2];
        V3 f = is_b ? V3{x, y, z} : V3{-x, -y, -z};
        f = inv_rotate_ti(f, tq.x, V3{tq.y, tq.z, tq.w});
        fx = add(fx, f.x); fy = add(fy, f.y); fz = add(fz, f.z);
        px = add(px, cp[c * 3]); py = add(py, cp[c * 3 + 1]); pz = add(pz, cp[c * 3 + 2]);
        cnt = add(cnt, 1.0f);
      }
    }
    if (cnt > 0.f) {
      px = fdiv(px, cnt); py = fdiv(py, cnt); pz = fdiv(pz, cnt);
    }
    float* of = out_f + ((size_t)e * Lc + t) * 3;
    float* op = out_p + ((size_t)e * Lc + t) * 3;
    of[0] = fx; of[1] = fy; of[2] = fz;
    op[0] = px; op[1] = py; op[2] = pz;
    counts[(size_t)e * Lc + t] = cnt;
  }
}

// ---------------------------------------------------------------------------------------------
// rotate_kernel: transform_by_quat(vec, q or conj(q)) for stand-alone getter calls
// ---------------------------------------------------------------------------------------------
__global__ void rotate_kernel(const float* __restrict__ vec, const float4* __restrict__ quat,
                              float* __restrict__ out, int n, int conjugate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 q = quat[i];
  const V3 qv = conjugate ? V3{-q.y, -q.z, -q.w} : V3{q.y, q.z, q.w};
  const V3 v = vec ? V3{vec[i * 3], vec[i * 3 + 1], vec[i * 3 + 2]} : V3{0.f, 0.f, -1.f};
  const V3 r = rotate(v, q.x, qv);
  out[i * 3] = r.x;
  out[i * 3 + 1] = r.y;
  out[i * 3 + 2] = r.z;
}

// ---------------------------------------------------------------------------------------------
// spawn_kernel: spawn pose of the reset envs in one launch (gfb_spawn_pose, include/gfb200.h).
// One thread per reset env; every scattered row belongs to exactly one thread.
// ---------------------------------------------------------------------------------------------
struct SpawnParams {
  gfb_spawn cfg;
  const int64_t* idx;
  int32_t n;
  const float* height_field;
  const float* u_x;
  const float* u_y;
  const float* u_rot[3];
  float* position_buffer;
  float* rot_buffer;
  float* quat_buffer;
  float* pos_out;
  float* quat_out;
};

__global__ void __launch_bounds__(128) spawn_kernel(const SpawnParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const gfb_spawn& c = p.cfg;
  const int64_t e = p.idx ? p.idx[i] : (int64_t)i;
  const Philox rng(c.rng_seed);
  const uint32_t e_lo = (uint32_t)e, e_hi = (uint32_t)((uint64_t)e >> 32);
  const uint32_t k_lo = (uint32_t)c.rng_counter, k_hi = (uint32_t)(c.rng_counter >> 32);

  float ux, uy;
  if (p.u_x && p.u_y) {
    ux = p.u_x[i];
    uy = p.u_y[i];
  } else {
    const uint4 r = rng(e_lo, e_hi ^ 0x53504157u, k_lo, k_hi);  // stream tag 'SPAW'
    ux = p.u_x ? p.u_x[i] : u01(r.x);
    uy = p.u_y ? p.u_y[i] : u01(r.y);
  }
  const float x = add(mul(ux, c.x_span), c.x_lo);
  const float y = add(mul(uy, c.y_span), c.y_lo);
  const float ground = p.height_field
                           ? terrain_height(x, y, c.terrain_bounds, c.height_field_rows, c.height_field_cols, p.height_field)
                           : c.flat_height;
  const float z = add(ground, c.height_offset);
  float* row = p.position_buffer + e * 3;
  row[0] = x;
  row[1] = y;
  row[2] = z;
  if (p.pos_out) {
    p.pos_out[(size_t)i * 3] = x;
    p.pos_out[(size_t)i * 3 + 1] = y;
    p.pos_out[(size_t)i * 3 + 2] = z;
  }
  if (!c.with_rotation) return;

  float ang[3];
  uint4 r{};
  bool drawn = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float* cell = p.rot_buffer + e * 3 + a;
    if (c.rot_mode[a] == GFB_SPAWN_ROT_DRAW) {
      float v;
      if (p.u_rot[a]) {
        v = p.u_rot[a][i];
      } else {
        if (!drawn) {
          r = rng(e_lo, e_hi ^ 0x53505254u, k_lo, k_hi);  // stream tag 'SPRT'
          drawn = true;
        }
        const uint32_t bits = a == 0 ? r.x : (a == 1 ? r.y : r.z);
        v = add(mul(u01(bits), sub(c.rot_hi[a], c.rot_lo[a])), c.rot_lo[a]);
      }
      *cell = v;
      ang[a] = v;
    } else {
      ang[a] = *cell;
    }
  }
  // xyz_to_quat (genesis.utils.geom; mdp/reset.py:63,194): extrinsic x-y-z, w first
  const float hx = mul(ang[0], 0.5f), hy = mul(ang[1], 0.5f), hz = mul(ang[2], 0.5f);
  const float cx = cosf(hx), cy = cosf(hy), cz = cosf(hz);
  const float sx = sinf(hx), sy = sinf(hy), sz = sinf(hz);
  const float qw = sub(mul(mul(cx, cy), cz), mul(mul(sx, sy), sz));
  const float qx = add(mul(mul(sx, cy), cz), mul(mul(cx, sy), sz));
  const float qy = sub(mul(mul(cx, sy), cz), mul(mul(sx, cy), sz));
  const float qz = add(mul(mul(cx, cy), sz), mul(mul(sx, sy), cz));
  const float4 q = make_float4(qw, qx, qy, qz);
  reinterpret_cast<float4*>(p.quat_buffer)[e] = q;
  if (p.quat_out) reinterpret_cast<float4*>(p.quat_out)[i] = q;
}

}  // namespace gfb
