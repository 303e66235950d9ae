// Device helpers for the fused manager-step kernels (sm_100a).
//
//  * rn-intrinsic arithmetic recipes that reproduce torch-CPU eager results bit for bit
//    (re-probed in DESIGN.md): every multiply/add separately rounded, never
//    left to nvcc's fmad contraction; cross products and short norms as the FMA chains the ATen
//    CPU kernels compile to.
//  * quaternion rotation following oracle/geom.py op for op.
//  * mbarrier + cp.async.bulk (TMA, non-tensor form) wrappers for slab loads/stores.
//  * Philox4x32-10 for the production (non-injected) random draws.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gfb {

// ---------------------------------------------------------------------------------------------
// exact-rounding arithmetic
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sq(float a) { return __fmul_rn(a, a); }

struct V3 {
  float x, y, z;
};

// torch.cross(a, b) on CPU evaluates each component as fma(a1, b2, -(a2*b1)).
__device__ __forceinline__ V3 cross(const V3& a, const V3& b) {
  V3 r;
  r.x = __fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y));
  r.y = __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z));
  r.z = __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x));
  return r;
}

// torch.norm(v, dim=-1) over 3 / 2 components: sqrt(fma(z,z,fma(y,y,x*x))).
__device__ __forceinline__ float norm3(float x, float y, float z) {
  return __fsqrt_rn(__fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x))));
}
__device__ __forceinline__ float norm2(float x, float y) {
  return __fsqrt_rn(__fmaf_rn(y, y, __fmul_rn(x, x)));
}

// oracle/geom.py transform_by_quat: t = cross(q_xyz, v) * 2; out = (v + w*t) + cross(q_xyz, t)
__device__ __forceinline__ V3 rotate(const V3& v, float w, const V3& q) {
  V3 t = cross(q, v);
  t.x = mul(t.x, 2.0f);
  t.y = mul(t.y, 2.0f);
  t.z = mul(t.z, 2.0f);
  V3 c = cross(q, t);
  V3 r;
  r.x = add(add(v.x, mul(w, t.x)), c.x);
  r.y = add(add(v.y, mul(w, t.y)), c.y);
  r.z = add(add(v.z, mul(w, t.z)), c.z);
  return r;
}

// oracle/geom.py ti_inv_transform_by_quat: q* = conj(q); u = q* x v; uu = q* x u; v + (w*u + uu)*2
__device__ __forceinline__ V3 inv_rotate_ti(const V3& v, float w, const V3& qv) {
  V3 q = {-qv.x, -qv.y, -qv.z};
  V3 u = cross(q, v);
  V3 uu = cross(q, u);
  V3 r;
  r.x = add(v.x, mul(add(mul(w, u.x), uu.x), 2.0f));
  r.y = add(v.y, mul(add(mul(w, u.y), uu.y), 2.0f));
  r.z = add(v.z, mul(add(mul(w, u.z), uu.z), 2.0f));
  return r;
}

__device__ __forceinline__ bool finite_f(float x) { return (__float_as_uint(x) & 0x7f800000u) != 0x7f800000u; }

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk async copies (TMA, linear form).  SASS: SYNCS.* / UBLKCP.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

//
// UNIFORM-DATAPATH RULE.  These instructions read their operands from uniform registers, which the 32
// lanes of a warp share.  Inside a lane-divergent branch (if (lane == 0) { ... }) the lanes that skipped
// the branch keep executing -- B200 interleaves the two paths when the guarded lane stalls on a fence /
// mbarrier instruction -- and may overwrite a uniform register the guarded lane is about to use.
// Observed (cuda-gdb, profiles/r2_01_mbarrier_init_clobber_evidence.txt): the low word of an mbarrier's
// init value replaced by an unrelated kernel parameter in ~15 % of the blocks => expected arrival count
// 0 => "Warp Illegal Instruction" at the barrier's second arm.  Therefore none of the wrappers below is
// ever called under a per-lane branch: the single lane is selected with a PREDICATE inside the asm
// statement (`pred`), the instruction stream stays convergent, there is no sibling path.
// tests/test_abi.py checks the SASS for it (tools/sass_lint.py).
__device__ __forceinline__ void mbar_init(bool pred, uint64_t* bar, uint32_t count) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %2, 0;\n"
      "@q mbarrier.init.shared::cta.b64 [%0], %1;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(count), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(bool pred, uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %2, 0;\n"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"((uint32_t)pred)
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
// global -> shared, completion counted on `bar`.  dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_load(bool pred, void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %4, 0;\n"
      "@q cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      "}\n" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "r"((uint32_t)pred)
      : "memory");
}
// shared -> global.
__device__ __forceinline__ void bulk_store(bool pred, void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %3, 0;\n"
      "@q cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
      "}\n" ::"l"(dst_gmem),
      "r"(smem_u32(src_smem)), "r"(bytes), "r"((uint32_t)pred)
      : "memory");
}
// (commit / wait are executed by every lane of the issuing warp: an empty group is a no-op)
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (before a bulk_store)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// TerrainManager.get_terrain_height (terrain_manager.py:100-166): normalise (x, y) to [-1, 1] by the
// terrain bounds with the reference's op sequence (sub, div, mul 2, sub 1), then bilinear
// grid_sample(padding=border, align_corners=True) on the (Hf, Wf) height field.
// bounds = {x_min, x_max, y_min, y_max}.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float terrain_height(float x, float y, const float* bounds, int Hf, int Wf,
                                                const float* __restrict__ hf) {
  const float xmin = bounds[0], xmax = bounds[1], ymin = bounds[2], ymax = bounds[3];
  const float gx = sub(mul(fdiv(sub(x, xmin), sub(xmax, xmin)), 2.0f), 1.0f);
  const float gy = sub(mul(fdiv(sub(y, ymin), sub(ymax, ymin)), 2.0f), 1.0f);
  float ix = mul(fdiv(add(gx, 1.0f), 2.0f), (float)(Wf - 1));
  float iy = mul(fdiv(add(gy, 1.0f), 2.0f), (float)(Hf - 1));
  ix = fminf(fmaxf(ix, 0.0f), (float)(Wf - 1));
  iy = fminf(fmaxf(iy, 0.0f), (float)(Hf - 1));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const int x1 = min(x0 + 1, Wf - 1), y1 = min(y0 + 1, Hf - 1);
  const float tx = sub(ix, fx0), ty = sub(iy, fy0);
  const float h00 = hf[y0 * Wf + x0], h01 = hf[y0 * Wf + x1];
  const float h10 = hf[y1 * Wf + x0], h11 = hf[y1 * Wf + x1];
  const float wx0 = sub(1.0f, tx), wy0 = sub(1.0f, ty);
  return add(add(mul(h00, mul(wx0, wy0)), mul(h01, mul(tx, wy0))),
             add(mul(h10, mul(wx0, ty)), mul(h11, mul(tx, ty))));
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------------------------
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ a, n1 = lo1, n2 = hi0 ^ c3 ^ b, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      a += 0x9E3779B9u;
      b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// uint32 -> [0, 1)
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

}  // namespace gfb
