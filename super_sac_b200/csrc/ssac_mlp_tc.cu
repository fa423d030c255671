// Ensemble MLP GEMMs on the 5th-generation tensor cores (impl = 2).  sm_100a only.
//
// Grouped tiles of C = opA(A) * opB(B) with fp32-accurate 3xTF32 arithmetic: every fp32 operand x is split into
// hi = tf32(x) and lo = x - hi, both staged in shared memory in canonical swizzled UMMA layouts, and each
// 8-deep k-step issues three tcgen05.mma.kind::tf32 (lo*hi, hi*lo, hi*hi) into one fp32 accumulator tile in
// tensor memory (TMEM).  The dropped lo*lo term is O(2^-22) relative, so results agree with an fp32 SGEMM to
// ~1e-6 (the reference runs true-fp32 SGEMM and north_star's tolerance is rtol 1e-4).
//
// Pipeline per CTA (one 128 x 128 output tile of one net, 256 threads):
//   all threads : global (L2) -> registers -> hi/lo split -> st.shared into stage s   (2 stages x 64 KB)
//                 (256 threads, four 16-byte loads per operand in flight per thread)
//   thread 0    : tcgen05.mma x (k-steps x 3) on stage s, tcgen05.commit -> mbarrier[s]  (frees the stage)
//   all threads : after the last commit, tcgen05.ld the 128 x N accumulator (thread = row), fused epilogue
//                 (bias, ReLU, ReLU-mask, extra gradient, accumulate), store.
// Operand layouts follow the contiguity of the source so that a 16-byte global load is a 16-byte shared store:
//   K-major  (source [rows][K] row-major):  addr(r,k) = (r/8)*1024 + (r%8)*128 + (((k/4) ^ (r%8)) * 16) + (k%4)*4
//             (SWIZZLE_128B: one 128-byte row per operand row and stage, 16-byte chunks XOR-swizzled with r%8, so
//              eight threads reading one contiguous 128-byte global row segment write eight distinct bank groups)
//   MN-major (source [K][cols] row-major):  addr(m,k) = (k/4)*2048 + (m/32)*512 + (k%4)*128
//                                                        + ((((m%32)/8) ^ (k%4)) * 32) + (m%8)*4
//             (SWIZZLE_128B_BASE32B: the only layout the tensor core accepts for MN-major 32-bit operands --
//              4 k-rows of 128 bytes per atom, 32-byte chunks XOR-swizzled with the k-row index)
#include "ssac_mlp.cuh"

namespace ssac {
namespace tc {

constexpr int TM = 128;      // MMA M (one CTA, cta_group::1)
constexpr int TN = 128;      // tile N (MMA N = 16..128, multiple of 16)
constexpr int TK = 32;       // k per pipeline stage (4 MMA k-steps of 8)
constexpr int kStages = 2;
constexpr int kProducerThreads = 256;  // warps 0-7 stage operands and run the epilogue (warps w, w+4 share TMEM lane quarter w%4)
constexpr int kThreads = kProducerThreads + 32;  // + warp 8: the MMA issuer
constexpr int kOperandBytes = TM * TK * 4;            // one hi or lo plane of one operand: 16 KB
constexpr int kStageBytes = 4 * kOperandBytes;        // A_hi, A_lo, B_hi, B_lo
constexpr int kSmemBytes = kStages * kStageBytes;     // 128 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory"); }

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

// shared-memory matrix descriptor, version 1 (Blackwell).  layout_type: 0 = none, 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout_type << 61);
}
// instruction descriptor: D fp32, A/B tf32, M = 128, N = n, majors: 0 = K-major, 1 = MN-major
__device__ __forceinline__ uint32_t instr_desc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
#ifdef SSAC_TRUNC_SPLIT
// experiment: the "hi" plane is the raw fp32 value (the tensor core ignores the low 13 mantissa bits of a tf32
// operand), lo = x - trunc_tf32(x)
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ void split_store(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t off, float4 v) {
  float4 l;
  l.x = v.x - tf32_trunc(v.x); l.y = v.y - tf32_trunc(v.y); l.z = v.z - tf32_trunc(v.z); l.w = v.w - tf32_trunc(v.w);
  *reinterpret_cast<float4*>(hi_plane + off) = v;
  *reinterpret_cast<float4*>(lo_plane + off) = l;
}
#else
__device__ __forceinline__ void split_store(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t off, float4 v) {
  float4 h, l;
  h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
  l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_plane + off) = h;
  *reinterpret_cast<float4*>(lo_plane + off) = l;
}
#endif

// ---- operand staging -------------------------------------------------------------------------------------------
// K-contiguous source [rows][K] (row-major, ld) -> K-major planes (SWIZZLE_128B).  Thread t owns the 4-k chunk
// c = t%8 of rows r = t/8 + 32i: a quarter-warp reads one contiguous 128-byte row segment and writes the eight
// swizzled 16-byte slots of one 128-byte shared row (coalesced and bank-conflict free).  All loads are issued
// before the first store so that four 16-byte requests per thread are in flight.
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
    split_store(hi, lo, (uint32_t)(r >> 3) * 1024u + r7 * 128u + (((uint32_t)c ^ r7) << 4), v[i]);
  }
}

// MN-contiguous source [K][cols] (row-major, ld) -> MN-major planes (SWIZZLE_128B_BASE32B).  Thread t owns the
// 4-column chunk mc = t%32 of rows k = t/32 + 8i: a quarter-warp reads 128 contiguous global bytes and writes one
// 128-byte shared row.
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
  const uint32_t mn_off = (uint32_t)(mc >> 3) * 512u + (uint32_t)(mc & 1) * 16u;
  const uint32_t chunk32 = (uint32_t)((mc & 7) >> 1);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kl = (t >> 5) + 8 * i;  // 0..31 inside the stage
    const uint32_t kr = (uint32_t)(kl & 3);
    split_store(hi, lo, (uint32_t)(kl >> 2) * 2048u + mn_off + kr * 128u + ((chunk32 ^ kr) << 5), v[i]);
  }
}

template <int LAYOUT>
__global__ void __launch_bounds__(kThreads, 1) grouped_gemm_tc_kernel(GemmP p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_full[kStages];    // producers -> MMA warp: stage s holds chunk kc
  __shared__ __align__(8) uint64_t bar_empty[kStages];   // tensor core -> producers: the MMAs reading stage s are done
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t tmem_base_sh;
  __shared__ float colsum_sh[8][TM];

  constexpr bool A_MN = (LAYOUT == L_TN);   // A given as [K][M]
  constexpr bool B_MN = (LAYOUT != L_NT);   // B given as [K][N] for NN / TN
  const int g = blockIdx.z;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int wg = p.b_index ? p.b_index[g] : g;
  const float* A = p.A + (int64_t)g * p.a_gs;
  const float* Bm = p.Bm + (int64_t)(LAYOUT == L_TN ? g : wg) * p.b_gs;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

  const int n_valid = min(TN, p.N - n0);
  const int n_mma = (n_valid + 15) & ~15;                 // MMA N: multiple of 16, 16..128
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < n_mma) tmem_cols <<= 1;

  if (warp == 0) tmem_alloc(&tmem_base_sh, tmem_cols);
  if (t == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], kProducerThreads);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_d = tmem_base_sh;
  const int nk = (p.K + TK - 1) / TK;

  if (warp == 8) {
    // ===== MMA issuer: one elected lane feeds the tensor core; completion is tracked by tcgen05.commit =====
    const uint32_t idesc = instr_desc(n_mma, A_MN ? 1 : 0, B_MN ? 1 : 0);
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc & 1, k0 = kc * TK;
      mbar_wait(&bar_full[s], (uint32_t)((kc >> 1) & 1));
      fence_after_sync();
      if (lane == 0) {
        const uint32_t st = smem_u32(smem) + (uint32_t)(s * kStageBytes);
        const uint32_t a_hi = st, a_lo = st + kOperandBytes, b_hi = st + 2 * kOperandBytes, b_lo = st + 3 * kOperandBytes;
        // K-major (SWIZZLE_128B)        : a k-step of 8 = 32 bytes further along the swizzled 128-byte rows,
        //                                 8-row groups SBO = 1024 apart (LBO unused)
        // MN-major (SWIZZLE_128B_BASE32B): a k-step of 8 = two 4-row k-groups SBO = 2048 apart,
        //                                 32-column groups LBO = 512 apart
        const uint32_t a_step = A_MN ? 4096u : 32u, b_step = B_MN ? 4096u : 32u;
        const uint32_t a_lbo = A_MN ? 512u : 16u, a_sbo = A_MN ? 2048u : 1024u, a_lt = A_MN ? 1u : 2u;
        const uint32_t b_lbo = B_MN ? 512u : 16u, b_sbo = B_MN ? 2048u : 1024u, b_lt = B_MN ? 1u : 2u;
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
    // ===== producers: global (L2) -> registers -> hi/lo split -> swizzled shared memory, one chunk ahead =====
    const bool a_vec = ((p.lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0) && (A_MN ? ((m0 & 3) == 0) : true);
    const bool b_vec = ((p.ldb & 3) == 0) && ((((uintptr_t)Bm) & 15) == 0) && (B_MN ? ((n0 & 3) == 0) : true);
    const bool do_colsum = (LAYOUT == L_TN) && p.colsum != nullptr && blockIdx.x == 0;
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
    float4 va[4], vb[4];
    if (nk > 0) {
      if (A_MN) load_mnmajor(va, A, p.lda, m0, p.M, 0, p.K, a_vec); else load_kmajor(va, A, p.lda, m0, p.M, 0, p.K, a_vec);
      if (B_MN) load_mnmajor(vb, Bm, p.ldb, n0, p.N, 0, p.K, b_vec); else load_kmajor(vb, Bm, p.ldb, n0, p.N, 0, p.K, b_vec);
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
    for (int kc = 0; kc < nk; ++kc) {
      const int s = kc & 1;
      uint8_t* st = smem + s * kStageBytes;
      uint8_t *a_hi = st, *a_lo = st + kOperandBytes, *b_hi = st + 2 * kOperandBytes, *b_lo = st + 3 * kOperandBytes;
      if (do_colsum) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { cs[0] += va[i].x; cs[1] += va[i].y; cs[2] += va[i].z; cs[3] += va[i].w; }
      }
      if (kc >= kStages) mbar_wait(&bar_empty[s], (uint32_t)(((kc >> 1) - 1) & 1));  // MMAs of chunk kc-2 done
      if (A_MN) store_mnmajor(a_hi, a_lo, va); else store_kmajor(a_hi, a_lo, va);
      if (B_MN) store_mnmajor(b_hi, b_lo, vb); else store_kmajor(b_hi, b_lo, vb);
      if (kc + 1 < nk) {   // next chunk's loads fly while the tensor core works on this one
        const int k1 = (kc + 1) * TK;
        if (A_MN) load_mnmajor(va, A, p.lda, m0, p.M, k1, p.K, a_vec); else load_kmajor(va, A, p.lda, m0, p.M, k1, p.K, a_vec);
        if (B_MN) load_mnmajor(vb, Bm, p.ldb, n0, p.N, k1, p.K, b_vec); else load_kmajor(vb, Bm, p.ldb, n0, p.N, k1, p.K, b_vec);
      }
      fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
      mbar_arrive(&bar_full[s]);
    }
    if (nk > 0) mbar_wait(&bar_done, 0);
    fence_after_sync();

    // ---- epilogue -----------------------------------------------------------------------------------------
    // phase 1: thread = accumulator row (TMEM lane quarter warp%4, column half warp/4) -> padded fp32 tile in
    //          shared memory (the stage buffers are free: every MMA has completed).
    // phase 2: coalesced pass, a warp per output row, fused bias / ReLU / extra / mask / accumulate.
    constexpr int kTilePitch = TN + 4;   // floats; +4 keeps 16-byte stores of 8 consecutive rows on distinct banks
    float* tile = reinterpret_cast<float*>(smem);
    {
      const int q = warp & 3, row = q * 32 + lane;
      for (int c0 = (warp >> 2) * 32; c0 < n_mma; c0 += 64) {
        float v[32];
        if (nk > 0) {
          tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);   // warp-collective
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        float* dst = tile + row * kTilePitch + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    producers_sync();
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
    if (LAYOUT == L_TN && p.colsum != nullptr && blockIdx.x == 0) {
      // thread t summed column chunk t%32 over its k rows (k = t/32 mod 8): combine the eight warps through smem
#pragma unroll
      for (int e = 0; e < 4; ++e) colsum_sh[warp][4 * lane + e] = cs[e];
      producers_sync();
      const int mm = m0 + t;
      if (t < TM && mm < p.M) {
        float tot = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) tot += colsum_sh[w8][t];
        float* out = p.colsum + (int64_t)wg * p.colsum_gs;
        out[mm] = p.accumulate ? (out[mm] + tot) : tot;
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

}  // namespace tc

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
  dim3 grid((p.N + tc::TN - 1) / tc::TN, (p.M + tc::TM - 1) / tc::TM, G);
  if (layout == L_NT) tc::grouped_gemm_tc_kernel<L_NT><<<grid, tc::kThreads, tc::kSmemBytes, s>>>(p);
  else if (layout == L_NN) tc::grouped_gemm_tc_kernel<L_NN><<<grid, tc::kThreads, tc::kSmemBytes, s>>>(p);
  else tc::grouped_gemm_tc_kernel<L_TN><<<grid, tc::kThreads, tc::kSmemBytes, s>>>(p);
  SSAC_CHECK_LAUNCH(what);
  return 0;
}

}  // namespace ssac
