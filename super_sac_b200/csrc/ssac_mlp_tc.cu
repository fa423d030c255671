// Ensemble MLP GEMMs on the 5th-generation tensor cores (impl = 2).  sm_100a only.
//
// Grouped tiles of C = opA(A) * opB(B) with fp32-accurate 3xTF32 arithmetic.  Every fp32 operand x is used as
//   hi = x as the tensor core sees it (kind::tf32 ignores the low 13 mantissa bits: measured on B200)
//   lo = x - trunc_tf32(x)   (exact in fp32)
// and each 8-deep k-step issues three tcgen05.mma.kind::tf32 (lo*hi, hi*lo, hi*hi) into one fp32 accumulator tile
// in tensor memory (TMEM).  The dropped lo*lo term is O(2^-20) relative (the reference runs true-fp32 SGEMM and
// north_star's tolerance is rtol 1e-4).
//
// Pipeline per CTA (one 128 x 128 output tile of one net; 10 warps, 3 stages of 64 KB):
//   warp 9      : TMA producer -- cp.async.bulk.tensor (3-D maps: k, row, net) drops the raw fp32 operand tiles
//                 straight into the swizzled UMMA layouts ("hi" planes), completion by mbarrier transaction bytes
//   warps 0-7   : element-wise lo pass smem -> smem (same offsets, no layout math); operands TMA cannot address
//                 (row pitch not a multiple of 16 bytes, e.g. the 23-wide first layer) are staged through registers
//   warp 8      : one elected lane issues the MMAs; tcgen05.commit releases the stage / signals the epilogue
//   warps 0-7   : epilogue -- tcgen05.ld (thread = accumulator row) -> padded smem tile -> coalesced pass with fused
//                 bias, ReLU, ReLU-mask, extra gradient, accumulate; bias gradients (column sums) ride along
// Operand layouts in shared memory (both what TMA writes and what the register path writes):
//   K-major  (source [rows][K] row-major):  SWIZZLE_128B          addr(r,k) = (r/8)*1024 + (r%8)*128
//                                                                            + (((k/4) ^ (r%8))*16) + (k%4)*4
//   MN-major (source [K][cols] row-major):  SWIZZLE_128B_BASE32B  addr(m,k) = (m/32)*4096 + k*128
//            (the only MN-major layout for 32-bit operands)                  + ((((m%32)/8) ^ (k%4))*32) + (m%8)*4
#include <cuda.h>

#include <algorithm>
#include <vector>

#include "ssac_tc_prims.cuh"

namespace ssac {
namespace tc {

#ifdef SSAC_TRACE
__device__ long long* g_trace = nullptr;
#define TRACE(slot)                                                                                   \
  do {                                                                                                \
    if (g_trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_trace[slot] = clock64(); \
  } while (0)
#else
#define TRACE(slot) do {} while (0)
#endif

struct GemmTC {
  GemmP p;
  CUtensorMap tmA, tmB, tmC;   // valid when a_tma / b_tma / c_tma
  int a_tma, b_tma, c_tma;
};

template <int LAYOUT>
__global__ void __launch_bounds__(kThreads, 1) grouped_gemm_tc_kernel(const __grid_constant__ GemmTC q) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_raw[kStages];     // TMA -> workers: the raw tiles of stage s have landed
  __shared__ __align__(8) uint64_t bar_full[kStages];    // workers -> MMA warp: lo planes written, stage s complete
  __shared__ __align__(8) uint64_t bar_empty[kStages];   // tensor core -> producers: the MMAs reading stage s are done
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_sh;
  __shared__ float bias_sh[TN];

  if (threadIdx.x == 0) TRACE(0);
  const GemmP& p = q.p;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle atoms need 1024-byte alignment
  constexpr bool A_MN = (LAYOUT == L_TN);   // A given as [K][M]
  constexpr bool B_MN = (LAYOUT != L_NT);   // B given as [K][N] for NN / TN
  const int g = blockIdx.z;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int wg = p.b_index ? p.b_index[g] : g;
  const int bz = (LAYOUT == L_TN) ? g : wg;   // group index of the B operand
  const float* A = p.A + (int64_t)g * p.a_gs;
  const float* Bm = p.Bm + (int64_t)bz * p.b_gs;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const bool a_tma = q.a_tma != 0, b_tma = q.b_tma != 0, any_tma = a_tma || b_tma;
  const int az_tma = p.a_gs ? g : 0, bz_tma = p.b_gs ? bz : 0;   // a shared operand has a single-group tensor map

  const int n_valid = min(TN, p.N - n0);
  const int n_mma = (n_valid + 15) & ~15;                 // MMA N: multiple of 16, 16..128
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_mma) tmem_cols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_base_sh, tmem_cols);
  if (t == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_raw[s], 1);
      mbar_init(&bar_full[s], kWorkerThreads);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();      // everything above is on-chip set-up; operands (and wg, below) may come from the previous kernel
  pdl_trigger();
  if (t < TN) bias_sh[t] = (p.bias && n0 + t < p.N) ? __ldg(p.bias + (int64_t)wg * p.bias_gs + n0 + t) : 0.f;
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = tmem_base_sh;
  const int nk = (p.K + TK - 1) / TK;
  if (threadIdx.x == 0) TRACE(1);

  if (warp == 9) {
    // ===== TMA producer ===========================================================================================
    if (any_tma && lane == 0) {
      const uint32_t bytes = (a_tma ? kPlaneBytes : 0) + (b_tma ? kPlaneBytes : 0);
      for (int kc = 0; kc < nk; ++kc) {
        const int s = kc % kStages, use = kc / kStages, k0 = kc * TK;
        if (kc >= kStages) mbar_wait(&bar_empty[s], (uint32_t)((use - 1) & 1));   // stage free again
        mbar_arrive_expect_tx(&bar_raw[s], bytes);
        const uint32_t st = smem_u32(smem) + (uint32_t)(s * kStageBytes);
        if (a_tma) {
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_3d(st + j * 4096, &q.tmA, &bar_raw[s], m0 + 32 * j, k0, az_tma);
          } else {
            tma_load_3d(st, &q.tmA, &bar_raw[s], k0, m0, az_tma);
          }
        }
        if (b_tma) {
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_3d(st + 2 * kPlaneBytes + j * 4096, &q.tmB, &bar_raw[s], n0 + 32 * j, k0, bz_tma);
          } else {
            tma_load_3d(st + 2 * kPlaneBytes, &q.tmB, &bar_raw[s], k0, n0, bz_tma);
          }
        }
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer: one elected lane feeds the tensor core; completion is tracked by tcgen05.commit ===========
    const uint32_t idesc = instr_desc(n_mma, A_MN ? 1 : 0, B_MN ? 1 : 0);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kStages, use = kc / kStages, k0 = kc * TK;
      mbar_wait(&bar_full[s], (uint32_t)(use & 1));
      fence_after_sync();
      if (elect_one()) {   // (one lane of the converged warp: no per-instruction uniformisation loops around the MMAs)
        const uint32_t st = smem_u32(smem) + (uint32_t)(s * kStageBytes);
        const uint32_t a_hi = st, a_lo = st + kPlaneBytes, b_hi = st + 2 * kPlaneBytes, b_lo = st + 3 * kPlaneBytes;
        // K-major (SWIZZLE_128B)        : a k-step of 8 = 32 bytes further along the swizzled 128-byte rows,
        //                                 8-row groups SBO = 1024 apart (LBO unused)
        // MN-major (SWIZZLE_128B_BASE32B): a k-step of 8 = two 4-row k-groups SBO = 512 apart (1024 bytes per step),
        //                                 32-column groups LBO = 4096 apart
        const uint32_t a_step = A_MN ? 1024u : 32u, b_step = B_MN ? 1024u : 32u;
        const uint32_t a_lbo = A_MN ? 4096u : 16u, a_sbo = A_MN ? 512u : 1024u, a_lt = A_MN ? 1u : 2u;
        const uint32_t b_lbo = B_MN ? 4096u : 16u, b_sbo = B_MN ? 512u : 1024u, b_lt = B_MN ? 1u : 2u;
        const int ksteps = min(TK / 8, (p.K - k0 + 7) / 8);
        for (int j = 0; j < ksteps; ++j) {
          const uint64_t dah = smem_desc(a_hi + j * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t dal = smem_desc(a_lo + j * a_step, a_lbo, a_sbo, a_lt);
          const uint64_t dbh = smem_desc(b_hi + j * b_step, b_lbo, b_sbo, b_lt);
          const uint64_t dbl = smem_desc(b_lo + j * b_step, b_lbo, b_sbo, b_lt);
          mma_tf32(tmem_d, dal, dbh, idesc, (kc | j) != 0);
          mma_tf32(tmem_d, dah, dbl, idesc, 1u);
          mma_tf32(tmem_d, dah, dbh, idesc, 1u);
        }
        mma_commit(&bar_empty[s]);
        if (kc == nk - 1) mma_commit(&bar_done);
      }
      __syncwarp();
    }
  } else {
    // ===== workers ================================================================================================
    const bool a_vec = ((p.lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0) && (A_MN ? ((m0 & 3) == 0) : true);
    const bool b_vec = ((p.ldb & 3) == 0) && ((((uintptr_t)Bm) & 15) == 0) && (B_MN ? ((n0 & 3) == 0) : true);
    const bool do_colsum = (LAYOUT == L_TN) && p.colsum != nullptr && blockIdx.x == 0;
    float cs[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) cs[e] = 0.f;
    float4 va[4], vb[4];
    if (nk > 0) {
      if (!a_tma) { if (A_MN) load_mnmajor(va, A, p.lda, m0, p.M, 0, p.K, a_vec); else load_kmajor(va, A, p.lda, m0, p.M, 0, p.K, a_vec); }
      if (!b_tma) { if (B_MN) load_mnmajor(vb, Bm, p.ldb, n0, p.N, 0, p.K, b_vec); else load_kmajor(vb, Bm, p.ldb, n0, p.N, 0, p.K, b_vec); }
    }
    // epilogue operands whose latency can hide behind the main loop
    const float* bias = p.bias ? p.bias + (int64_t)wg * p.bias_gs : nullptr;
    const int nc = n0 + 4 * lane;                       // this lane's first output column in the coalesced pass
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (bias) {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (nc + e < p.N) bv[e] = __ldg(bias + nc + e);
    }
    // gate (see GemmP::a_gate): the four floats of a 16-byte chunk are consecutive along A's contiguous dimension; the
    // chunk's position follows from the swizzle formulas at the top of this file.  MN-major A: the gate index is the m
    // coordinate, the same for every stage (loaded once); K-major A: it is the k coordinate -- the values of stage kc + 1
    // are requested while stage kc is converted, so their latency stays off the stage's critical path.
    const float* gate = (a_tma && p.a_gate) ? p.a_gate + (int64_t)wg * p.a_gate_gs : nullptr;
    float4 gv[4];
    auto load_gate = [&](int kc) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = (uint32_t)(t + kWorkerThreads * i) * 16u;
        int j0, lim;
        if (A_MN) {   // (m/32)*4096 + k*128 + (((m%32)/8 ^ k%4)*32) + (m%8)*4
          const uint32_t kk = (off >> 7) & 31u;
          j0 = m0 + (int)(32u * (off >> 12) + 8u * (((off >> 5) & 3u) ^ (kk & 3u)) + 4u * ((off >> 4) & 1u));
          lim = p.M;
        } else {      // r*128 + ((k/4 ^ r%8)*16) + (k%4)*4
          const uint32_t r = off >> 7;
          j0 = kc * TK + (int)(4u * (((off >> 4) & 7u) ^ (r & 7u)));
          lim = p.K;
        }
        gv[i] = (j0 + 3 < lim) ? __ldg(reinterpret_cast<const float4*>(gate + j0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if (gate && nk > 0) load_gate(0);
    const bool kscale = (LAYOUT == L_TN) && a_tma && p.a_kscale != nullptr;
    auto load_ksc = [&](int kc) {
      const int kk = kc * TK + (t >> 3);
      return kk < p.K ? __ldg(p.a_kscale + (int64_t)g * p.a_kscale_gs + kk) : 0.f;
    };
    float ksc_next = (kscale && nk > 0) ? load_ksc(0) : 1.f;
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc % kStages, use = kc / kStages;
      uint8_t* st = smem + s * kStageBytes;
      uint8_t *a_hi = st, *a_lo = st + kPlaneBytes, *b_hi = st + 2 * kPlaneBytes, *b_lo = st + 3 * kPlaneBytes;
      if (t == 0 && kc < 8) TRACE(2 + 3 * kc);
      if (any_tma) mbar_wait(&bar_raw[s], (uint32_t)(use & 1));                   // raw tiles landed (=> stage was free)
      else if (kc >= kStages) mbar_wait(&bar_empty[s], (uint32_t)((use - 1) & 1));  // no TMA operand: wait for the MMAs
      if (t == 0 && kc < 8) TRACE(3 + 3 * kc);
      if (a_tma) {
        // lo pass: same (swizzled) offsets in and out, 4 x 16 bytes per thread.  In the MN-major layout all four
        // chunks of a thread sit on k row t/8 of the stage, so an optional per-k scale (the TD-error seed of the
        // split critic backward, mlp_backward_post) is one load per thread and stage.
        const float ksc = ksc_next;
        if (kscale && kc + 1 < nk) ksc_next = load_ksc(kc + 1);   // (requested a stage ahead, like the gate values)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = (uint32_t)(t + kWorkerThreads * i) * 16u;
          float4 v = *reinterpret_cast<const float4*>(a_hi + off);
          if (gate) {
            v.x = v.x > 0.f ? gv[i].x : 0.f; v.y = v.y > 0.f ? gv[i].y : 0.f;
            v.z = v.z > 0.f ? gv[i].z : 0.f; v.w = v.w > 0.f ? gv[i].w : 0.f;
          }
          if (kscale) { v.x *= ksc; v.y *= ksc; v.z *= ksc; v.w *= ksc; }
          if (kscale || gate) *reinterpret_cast<float4*>(a_hi + off) = v;
          *reinterpret_cast<float4*>(a_lo + off) = lo4(v);
          if (do_colsum) { cs[4 * i + 0] += v.x; cs[4 * i + 1] += v.y; cs[4 * i + 2] += v.z; cs[4 * i + 3] += v.w; }
        }
        if (gate && !A_MN && kc + 1 < nk) load_gate(kc + 1);
      } else {
        if (do_colsum) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { cs[0] += va[i].x; cs[1] += va[i].y; cs[2] += va[i].z; cs[3] += va[i].w; }
        }
        if (A_MN) store_mnmajor(a_hi, a_lo, va); else store_kmajor(a_hi, a_lo, va);
      }
      if (b_tma) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t off = (uint32_t)(t + kWorkerThreads * i) * 16u;
          *reinterpret_cast<float4*>(b_lo + off) = lo4(*reinterpret_cast<const float4*>(b_hi + off));
        }
      } else {
        if (B_MN) store_mnmajor(b_hi, b_lo, vb); else store_kmajor(b_hi, b_lo, vb);
      }
      if (kc + 1 < nk) {   // register-staged operands: next chunk's loads fly while the tensor core works
        const int k1 = (kc + 1) * TK;
        if (!a_tma) { if (A_MN) load_mnmajor(va, A, p.lda, m0, p.M, k1, p.K, a_vec); else load_kmajor(va, A, p.lda, m0, p.M, k1, p.K, a_vec); }
        if (!b_tma) { if (B_MN) load_mnmajor(vb, Bm, p.ldb, n0, p.N, k1, p.K, b_vec); else load_kmajor(vb, Bm, p.ldb, n0, p.N, k1, p.K, b_vec); }
      }
      fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&bar_full[s]);
      if (t == 0 && kc < 8) TRACE(4 + 3 * kc);
    }
    if (t == 0) TRACE(30);
    if (nk > 0) mbar_wait(&bar_done, 0);
    fence_after_sync();
    if (t == 0) TRACE(31);

    if (q.c_tma) {
      // ---- epilogue A (no mask / extra / accumulate): bias + ReLU in registers (thread = accumulator row), tile
      // written in the SWIZZLE_128B layout of a [32 col x 128 row] TMA box per column block, then ONE thread issues
      // cp.async.bulk.tensor stores; rows / columns beyond M / N are clipped by the tensor map.
      const int qd = warp & 3, row = qd * 32 + lane;
      const uint32_t r7 = (uint32_t)(row & 7);
      for (int c0 = (warp >> 2) * 32; c0 < n_mma; c0 += 64) {
        float v[32];
        if (nk > 0) {
          tmem_ld32(tmem_d + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);   // warp-collective
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        uint8_t* box = smem + (c0 >> 5) * kPlaneBytes + (uint32_t)(row >> 3) * 1024u + r7 * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 o;
          o.x = v[4 * c + 0] + bias_sh[c0 + 4 * c + 0]; o.y = v[4 * c + 1] + bias_sh[c0 + 4 * c + 1];
          o.z = v[4 * c + 2] + bias_sh[c0 + 4 * c + 2]; o.w = v[4 * c + 3] + bias_sh[c0 + 4 * c + 3];
          if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
          *reinterpret_cast<float4*>(box + (((uint32_t)c ^ r7) << 4)) = o;
        }
      }
      if (do_colsum) {
        float* part = reinterpret_cast<float*>(smem + 96 * 1024);
        if (a_tma) {
          const int kr = (t >> 3) & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int mbase = 32 * i + ((((t & 7) >> 1) ^ kr) << 3) + ((t & 1) << 2);
#pragma unroll
            for (int e = 0; e < 4; ++e) part[(t >> 3) * TM + mbase + e] = cs[4 * i + e];
          }
        } else {
          const int mbase = 4 * (t & 31);
#pragma unroll
          for (int e = 0; e < 4; ++e) part[(t >> 5) * TM + mbase + e] = cs[e];
        }
      }
      fence_async_smem();
      workers_sync();
      if (t == 0) {
        const int cz = (LAYOUT == L_TN) ? wg : g;
        for (int c0 = 0; c0 < n_mma; c0 += 32)
          if (n0 + c0 < p.N) tma_store_3d(&q.tmC, smem_u32(smem) + (uint32_t)((c0 >> 5) * kPlaneBytes), n0 + c0, m0, p.c_gs ? cz : 0);
        tma_store_commit_and_wait_read();
      }
      if (do_colsum) {
        const float* part = reinterpret_cast<const float*>(smem + 96 * 1024);
        const int n_slots = a_tma ? 32 : 8;
        const int mm = m0 + t;
        if (t < TM && mm < p.M) {
          float tot = 0.f;
          for (int sl = 0; sl < n_slots; ++sl) tot += part[sl * TM + t];
          float* out = p.colsum + (int64_t)wg * p.colsum_gs;
          out[mm] = p.accumulate ? (out[mm] + tot) : tot;
        }
      }
    } else {
    // ---- epilogue -----------------------------------------------------------------------------------------
    // phase 1: thread = accumulator row (TMEM lane quarter warp%4, column half warp/4) -> padded fp32 tile in
    //          shared memory (the stage buffers are free: every MMA has completed).
    // phase 2: coalesced pass, a warp per output row, fused bias / ReLU / extra / mask / accumulate.
    constexpr int kTilePitch = TN + 4;   // floats; +4 keeps 16-byte stores of 8 consecutive rows on distinct banks
    float* tile = reinterpret_cast<float*>(smem);
    {
      const int qd = warp & 3, row = qd * 32 + lane;
      for (int c0 = (warp >> 2) * 32; c0 < n_mma; c0 += 64) {
        float v[32];
        if (nk > 0) {
          tmem_ld32(tmem_d + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, v);   // warp-collective
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        float* dst = tile + row * kTilePitch + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    // bias gradients: every (k-row slot, column) partial has exactly one owner thread; they are written to a
    // [slots][128] scratch above the output tile and summed in a fixed order (bit-reproducible, no atomics)
    float* part = reinterpret_cast<float*>(smem + 96 * 1024);
    const int n_slots = a_tma ? 32 : 8;
    if (do_colsum) {
      if (a_tma) {
        // chunk i of thread t sits at byte offset 16*(t + 256 i): m-group i, k row t/8, physical 32-byte chunk (t%8)/2
        const int kr = (t >> 3) & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int mbase = 32 * i + ((((t & 7) >> 1) ^ kr) << 3) + ((t & 1) << 2);
#pragma unroll
          for (int e = 0; e < 4; ++e) part[(t >> 3) * TM + mbase + e] = cs[4 * i + e];
        }
      } else {
        const int mbase = 4 * (t & 31);
#pragma unroll
        for (int e = 0; e < 4; ++e) part[(t >> 5) * TM + mbase + e] = cs[e];
      }
    }
    workers_sync();
    if (t == 0) TRACE(32);
    {
      float* C = p.C + (int64_t)(LAYOUT == L_TN ? wg : g) * p.c_gs;
      const float* mask = p.mask ? p.mask + (int64_t)g * p.mask_gs : nullptr;
      const float* extra = p.extra ? p.extra + (int64_t)g * p.extra_gs : nullptr;
      const bool full4 = nc + 3 < p.N;
      const bool c_vec = ((p.ldc & 3) == 0) && ((((uintptr_t)C) & 15) == 0) && full4;
      const bool x_vec = extra && ((p.ldextra & 3) == 0) && ((((uintptr_t)extra) & 15) == 0) && full4;
      const bool m_vec = mask && ((p.ldmask & 3) == 0) && ((((uintptr_t)mask) & 15) == 0) && full4;
      if (4 * lane < n_mma && nc < p.N) {
        // 16 rows per warp, 4 at a time: all loads of a group are issued before the first store
        for (int r0 = warp; r0 < TM; r0 += 32) {
          float x[4][4], ev[4][4], mv[4][4], cv[4][4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + 8 * u, m = m0 + r;
            const float4 t4 = *reinterpret_cast<const float4*>(tile + r * kTilePitch + 4 * lane);
            x[u][0] = t4.x; x[u][1] = t4.y; x[u][2] = t4.z; x[u][3] = t4.w;
#pragma unroll
            for (int e = 0; e < 4; ++e) { ev[u][e] = 0.f; mv[u][e] = 1.f; cv[u][e] = 0.f; }
            if (m < p.M) {
              if (extra) {
                const float* ep = extra + (int64_t)m * p.ldextra + nc;
                if (x_vec) { const float4 q4 = *reinterpret_cast<const float4*>(ep); ev[u][0] = q4.x; ev[u][1] = q4.y; ev[u][2] = q4.z; ev[u][3] = q4.w; }
                else { for (int e = 0; e < 4; ++e) if (nc + e < p.N) ev[u][e] = ep[e]; }
              }
              if (mask) {
                const float* mp = mask + (int64_t)m * p.ldmask + nc;
                if (m_vec) { const float4 q4 = *reinterpret_cast<const float4*>(mp); mv[u][0] = q4.x; mv[u][1] = q4.y; mv[u][2] = q4.z; mv[u][3] = q4.w; }
                else { for (int e = 0; e < 4; ++e) if (nc + e < p.N) mv[u][e] = mp[e]; }
              }
              if (p.accumulate) {
                const float* cp = C + (int64_t)m * p.ldc + nc;
                if (c_vec) { const float4 q4 = *reinterpret_cast<const float4*>(cp); cv[u][0] = q4.x; cv[u][1] = q4.y; cv[u][2] = q4.z; cv[u][3] = q4.w; }
                else { for (int e = 0; e < 4; ++e) if (nc + e < p.N) cv[u][e] = cp[e]; }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = r0 + 8 * u, m = m0 + r;
            if (m >= p.M) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float y = x[u][e] + bv[e];
              if (p.relu) y = fmaxf(y, 0.f);
              y += p.extra_scale * ev[u][e];
              y = mv[u][e] > 0.f ? y : 0.f;
              x[u][e] = y + cv[u][e];
            }
            float* cp = C + (int64_t)m * p.ldc + nc;
            if (c_vec) {
              *reinterpret_cast<float4*>(cp) = make_float4(x[u][0], x[u][1], x[u][2], x[u][3]);
            } else {
              for (int e = 0; e < 4; ++e)
                if (nc + e < p.N) cp[e] = x[u][e];
            }
          }
        }
      }
    }
    if (do_colsum) {
      const int mm = m0 + t;
      if (t < TM && mm < p.M) {
        float tot = 0.f;
        for (int sl = 0; sl < n_slots; ++sl) tot += part[sl * TM + t];
        float* out = p.colsum + (int64_t)wg * p.colsum_gs;
        out[mm] = p.accumulate ? (out[mm] + tot) : tot;
      }
    }
    }  // epilogue B
  }
  if (t == 0) TRACE(33);
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
  if (t == 0) TRACE(34);
}

}  // namespace tc

// ---- host: tensor maps ---------------------------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// cuMemGetAddressRange: the [base, base + size) of the allocation (cudaMalloc block / torch allocator segment) holding p
typedef CUresult (*AddrRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
AddrRangeFn addr_range_fn() {
  static AddrRangeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<AddrRangeFn>(ptr);
  }
  return fn;
}

struct MapKey {
  const void* base; int64_t ld, gs; int inner, outer; bool mn; int64_t ngroups;
  bool operator==(const MapKey& o) const {
    return base == o.base && ld == o.ld && gs == o.gs && inner == o.inner && outer == o.outer && mn == o.mn && ngroups == o.ngroups;
  }
};
struct MapEntry { MapKey key; CUtensorMap map; };
std::vector<MapEntry>& map_cache() {
  static thread_local std::vector<MapEntry> cache;
  return cache;
}

}  // namespace

namespace tc {
// 3-D map (inner, outer, group) over a row-major fp32 matrix stack.  K-major operands: inner = k, outer = rows,
// box 32 x 128, SWIZZLE_128B.  MN-major operands: inner = cols, outer = k, box 32 x 32, SWIZZLE_128B_ATOM_32B.
//
// MEASURED on B200 (tools/probes/tma_tail_probe.cu, profiles/r2_03_tma_tail_probe.log): when a box hangs over the
// tensor's bounds (rows >= outer or columns >= inner, zero-filled as they should be) the TMA unit still touches global
// addresses up to ~64 KB behind the box -- and FAULTS if they are unmapped -- whenever the tensor map DECLARES an extent
// that reaches there.  A group dimension declared "large enough" (this code used 65536) over a stack that ends at the
// tail of its allocation did exactly that.  With the declared extent inside mapped memory no configuration faulted, at
// any distance from the end of the mapping; boxes that do not hang over never fault either.  Hence:
//   * the group count of a map is clamped to the groups that fit the allocation holding `base` (cuMemGetAddressRange),
//     so the declared extent never leaves it; `groups` (the groups this launch addresses without a net subset) must fit;
//   * if the allocation cannot be queried the map is only handed out for boxes that cannot hang over.
// Every TMA operand has a register-staged path for a refused map.
bool make_map(const float* base, int64_t ld, int64_t gs, int inner, int outer, bool mn, CUtensorMap* out, int groups) {
  if (!encode_fn()) return false;
  if ((((uintptr_t)base) & 15) != 0 || (ld & 3) != 0 || ld <= 0 || (gs & 3) != 0 || inner <= 0 || outer <= 0) return false;
  const int64_t box_rows = mn ? 32 : 128;
  int64_t ngroups = gs > 0 ? 65536 : 1;
  bool known = false;
  if (AddrRangeFn range = addr_range_fn()) {
    CUdeviceptr abase = 0;
    size_t asize = 0;
    if (range(&abase, &asize, (CUdeviceptr)(uintptr_t)base) == CUDA_SUCCESS && asize > 0) {
      const int64_t room = (int64_t)((uintptr_t)abase + asize - (uintptr_t)base) / 4;      // floats from base to the end
      const int64_t one = (int64_t)(outer - 1) * ld + inner;                                // footprint of one group
      if (room < one) return false;
      if (gs > 0) ngroups = std::min<int64_t>(ngroups, (room - one) / gs + 1);
      known = true;
    }
  }
  static const bool dbg = getenv("SSAC_DEBUG_MAPS") != nullptr;
  if (dbg)
    fprintf(stderr, "[make_map] base %p ld %lld gs %lld inner %d outer %d mn %d groups %d: declared groups %lld (%s)\n",
            (const void*)base, (long long)ld, (long long)gs, inner, outer, (int)mn, groups, (long long)ngroups,
            known ? "allocation known" : "allocation unknown");
  if (!known && (inner % 32 != 0 || outer % box_rows != 0)) return false;
  if (gs > 0 && ngroups < groups) return false;
  MapKey key{base, ld, gs, inner, outer, mn, ngroups};
  auto& cache = map_cache();
  for (auto& e : cache)
    if (e.key == key) { *out = e.map; return true; }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)outer, (cuuint64_t)ngroups};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)(gs > 0 ? gs : (int64_t)outer * ld) * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() >= 256) cache.clear();
  cache.push_back(MapEntry{key, *out});
  return true;
}
bool make_map2d(const float* base, int64_t ld, int inner, int64_t rows, int box_rows, bool mn, CUtensorMap* out) {
  if (!encode_fn()) return false;
  if ((((uintptr_t)base) & 15) != 0 || (ld & 3) != 0 || ld < inner || inner <= 0 || rows <= 0 || box_rows <= 0 || box_rows > 256) return false;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
}  // namespace tc
using tc::make_map;

#ifdef SSAC_TRACE
extern "C" int ssac_debug_set_trace(long long* dev_ptr) { return (int)cudaMemcpyToSymbol(tc::g_trace, &dev_ptr, sizeof(dev_ptr)); }
#endif

static bool g_tma_enabled = true;
bool tc::tma_enabled() { return g_tma_enabled; }
extern "C" int ssac_set_tma_enabled(int on) {
  g_tma_enabled = on != 0;
  return 0;
}

int launch_gemm_tc(int layout, const GemmP& p, int G, cudaStream_t s, const char* what) {
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[layout]) {
    cudaError_t e;
    if (layout == L_NT) e = cudaFuncSetAttribute(tc::grouped_gemm_tc_kernel<L_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    else if (layout == L_NN) e = cudaFuncSetAttribute(tc::grouped_gemm_tc_kernel<L_NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    else e = cudaFuncSetAttribute(tc::grouped_gemm_tc_kernel<L_TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    if (e != cudaSuccess) {
      set_error(std::string(what) + " (smem attribute): " + cudaGetErrorString(e));
      return (int)e;
    }
    attr_set[layout] = true;
  }
  tc::GemmTC q;
  q.p = p;
  q.a_tma = q.b_tma = q.c_tma = 0;
  if (g_tma_enabled && p.K > 0) {
    const bool a_mn = layout == L_TN, b_mn = layout != L_NT;
    // group strides are baked into the maps; base pointers are the group-0 matrices
    if (a_mn) q.a_tma = make_map(p.A, p.lda, p.a_gs, p.M, p.K, true, &q.tmA, G);
    else q.a_tma = make_map(p.A, p.lda, p.a_gs, p.K, p.M, false, &q.tmA, G);
    if (b_mn) q.b_tma = make_map(p.Bm, p.ldb, p.b_gs, p.N, p.K, true, &q.tmB, G);
    else q.b_tma = make_map(p.Bm, p.ldb, p.b_gs, p.K, p.N, false, &q.tmB, G);
    if (!p.mask && !p.extra && !p.accumulate) q.c_tma = make_map(p.C, p.ldc, p.c_gs, p.N, p.M, false, &q.tmC, G);
  }
  if (p.a_kscale && !(layout == L_TN && q.a_tma))
    return fail(SSAC_E_UNSUPPORTED, "tcgen05 GEMM: a per-k scale needs a TMA-addressable transposed A operand");
  if (p.a_gate && !(q.a_tma && (layout == L_TN ? p.M : p.K) % 4 == 0 && (((uintptr_t)p.a_gate | (uintptr_t)(p.a_gate_gs * 4)) & 15) == 0))
    return fail(SSAC_E_UNSUPPORTED, "tcgen05 GEMM: a gated A operand needs a TMA-addressable A and 16-byte aligned gate rows");
  dim3 grid((p.N + tc::TN - 1) / tc::TN, (p.M + tc::TM - 1) / tc::TM, G);
  if (p.pdl) {
    if (layout == L_NT) launch_pdl(tc::grouped_gemm_tc_kernel<L_NT>, grid, dim3(tc::kThreads), tc::kSmemBytes, s, q);
    else if (layout == L_NN) launch_pdl(tc::grouped_gemm_tc_kernel<L_NN>, grid, dim3(tc::kThreads), tc::kSmemBytes, s, q);
    else launch_pdl(tc::grouped_gemm_tc_kernel<L_TN>, grid, dim3(tc::kThreads), tc::kSmemBytes, s, q);
  } else {
    if (layout == L_NT) tc::grouped_gemm_tc_kernel<L_NT><<<grid, tc::kThreads, tc::kSmemBytes, s>>>(q);
    else if (layout == L_NN) tc::grouped_gemm_tc_kernel<L_NN><<<grid, tc::kThreads, tc::kSmemBytes, s>>>(q);
    else tc::grouped_gemm_tc_kernel<L_TN><<<grid, tc::kThreads, tc::kSmemBytes, s>>>(q);
  }
  SSAC_CHECK_LAUNCH(what);
  return 0;
}

}  // namespace ssac
