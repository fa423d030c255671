// tcgen05 / TMEM / TMA / mbarrier primitives shared by the tensor-core kernels (sm_100a only): inline PTX wrappers,
// the 3xTF32 operand split and the register-staging helpers for operands TMA cannot address.
#pragma once
#include <cuda.h>

#include "ssac_mlp.cuh"

namespace ssac {
namespace tc {

constexpr int TM = 128;      // MMA M (one CTA, cta_group::1)
constexpr int TN = 128;      // tile N (MMA N = 16..128, multiple of 16)
constexpr int TK = 32;       // k per pipeline stage (4 MMA k-steps of 8) = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kWorkerThreads = 256;               // warps 0-7
constexpr int kThreads = kWorkerThreads + 64;     // + warp 8 (MMA issuer) + warp 9 (TMA producer)
constexpr int kPlaneBytes = TM * TK * 4;          // one hi or lo plane of one operand: 16 KB
constexpr int kStageBytes = 4 * kPlaneBytes;      // A_hi, A_lo, B_hi, B_lo
constexpr int kSmemBytes = kStages * kStageBytes + 1024;   // + slack for 1024-byte alignment of the swizzle atoms

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 50000000u) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkerThreads) : "memory"); }

// ---- TMA ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 2-D variants (conv encoder: (channels, pixels) views of NHWC activations)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- tensor memory / tcgen05 -------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a converged warp (elect.sync): lets the compiler keep tcgen05 operands in uniform registers without a
// per-instruction uniformisation loop, which `if (lane == 0)` forces
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem], tf32 inputs, fp32 accumulate.  Issued by ONE thread for the whole CTA.
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// mbarrier arrives when every tcgen05.mma issued so far by this thread has completed (implies before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 consecutive fp32 accumulator columns of this thread's TMEM lane -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, version 1 (Blackwell).  layout_type: 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// instruction descriptor: D fp32, A/B tf32, M = 128, N = n, majors: 0 = K-major, 1 = MN-major
__device__ __forceinline__ uint32_t instr_desc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float4 lo4(float4 v) { return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)); }

// ---- register staging for operands TMA cannot address ------------------------------------------------------------
// K-contiguous source [rows][K] -> K-major planes.  Thread t owns the 4-k chunk c = t%8 of rows r = t/8 + 32i.
__device__ __forceinline__ void load_kmajor(float4 (&v)[4], const float* __restrict__ src, int64_t ld, int row0,
                                            int nrows, int k0, int K, bool vec_ok) {
  const int t = threadIdx.x, c = t & 7;
  const int kc = k0 + 4 * c;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (t >> 3) + 32 * i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < nrows) {
      const float* rp = src + (int64_t)(row0 + r) * ld + kc;
      if (vec_ok && kc + 3 < K) {
        v[i] = __ldg(reinterpret_cast<const float4*>(rp));
      } else {
        if (kc + 0 < K) v[i].x = __ldg(rp + 0);
        if (kc + 1 < K) v[i].y = __ldg(rp + 1);
        if (kc + 2 < K) v[i].z = __ldg(rp + 2);
        if (kc + 3 < K) v[i].w = __ldg(rp + 3);
      }
    }
  }
}
__device__ __forceinline__ void store_kmajor(uint8_t* hi, uint8_t* lo, const float4 (&v)[4]) {
  const int t = threadIdx.x, c = t & 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (t >> 3) + 32 * i;
    const uint32_t r7 = (uint32_t)(r & 7);
    const uint32_t off = (uint32_t)(r >> 3) * 1024u + r7 * 128u + (((uint32_t)c ^ r7) << 4);
    *reinterpret_cast<float4*>(hi + off) = v[i];
    *reinterpret_cast<float4*>(lo + off) = lo4(v[i]);
  }
}
// MN-contiguous source [K][cols] -> MN-major planes.  Thread t owns the 4-column chunk mc = t%32 of rows k = t/32 + 8i.
__device__ __forceinline__ void load_mnmajor(float4 (&v)[4], const float* __restrict__ src, int64_t ld, int col0,
                                             int ncols, int k0, int K, bool vec_ok) {
  const int t = threadIdx.x, mc = t & 31, c = col0 + 4 * mc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = k0 + (t >> 5) + 8 * i;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
      const float* rp = src + (int64_t)k * ld + c;
      if (vec_ok && c + 3 < ncols) {
        v[i] = __ldg(reinterpret_cast<const float4*>(rp));
      } else {
        if (c + 0 < ncols) v[i].x = __ldg(rp + 0);
        if (c + 1 < ncols) v[i].y = __ldg(rp + 1);
        if (c + 2 < ncols) v[i].z = __ldg(rp + 2);
        if (c + 3 < ncols) v[i].w = __ldg(rp + 3);
      }
    }
  }
}
__device__ __forceinline__ void store_mnmajor(uint8_t* hi, uint8_t* lo, const float4 (&v)[4]) {
  const int t = threadIdx.x, mc = t & 31;
  const uint32_t mn_off = (uint32_t)(mc >> 3) * 4096u + (uint32_t)(mc & 1) * 16u;
  const uint32_t chunk32 = (uint32_t)((mc & 7) >> 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t kl = (uint32_t)((t >> 5) + 8 * i);  // 0..31 inside the stage
    const uint32_t off = mn_off + kl * 128u + ((chunk32 ^ (kl & 3u)) << 5);
    *reinterpret_cast<float4*>(hi + off) = v[i];
    *reinterpret_cast<float4*>(lo + off) = lo4(v[i]);
  }
}

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane -> registers
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// issue-only variants: the registers are valid after tmem_wait_ld() (several loads can be in flight)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// host side (ssac_mlp_tc.cu): cached 3-D tensor maps (inner, outer, group) over row-major fp32 matrix stacks
bool make_map(const float* base, int64_t ld, int64_t gs, int inner, int outer, bool mn, CUtensorMap* out, int groups = 1);
bool tma_enabled();
// 2-D map (inner, rows) over a dense row-major fp32 matrix whose rows are `ld` floats apart; box = 32 x box_rows;
// mn = false: SWIZZLE_128B (K-major operand / plain tile), true: SWIZZLE_128B_ATOM_32B (MN-major operand).  The declared
// extent is exactly inner x rows (boxes hanging over it are zero-filled / clipped; see make_map for why it must not be more).
bool make_map2d(const float* base, int64_t ld, int inner, int64_t rows, int box_rows, bool mn, CUtensorMap* out);

}  // namespace tc
}  // namespace ssac
