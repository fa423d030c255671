// Row-local MLP chains on the CUDA cores, exact fp32: the latency-bound single-net forwards of the update step.
//
// The TD-target path of one member -- a1, logp = pi(s1), then the M target critics of the REDQ subset on (s1, a1)
// (learning_utils.py:314-338, agent.py:22-40) -- is 0.11 GFLOP on B = 256 rows: as two tensor-core launches its grids are
// 8 and 16 CTAs and each launch pays ~15 us of serial latency (TMEM allocation, a K = 256 pipeline of eight chunks,
// epilogue).  Every step of that chain is ROW-LOCAL, so here ONE launch of ceil(B/8) clusters x 4 CTAs (128 CTAs for
// B = 256) walks the whole chain for 8 batch rows per cluster:
//
//   per net  : CTA r of the cluster owns hidden units [r*H/4, (r+1)*H/4) of both hidden layers
//     layer 1: K = D <= 64, weights slice staged in shared memory, thread = (unit, row group)
//              -> the 8 x H/4 activations are broadcast into all four CTAs' shared memory (DSMEM, 16-byte stores)
//     layer 2: thread = (unit pair, k split): its 2 x 32 weights come straight from L2 into registers (sixteen 16-byte
//              loads in flight per thread, each weight is used for all 8 rows), activations are warp-broadcast 16-byte
//              shared-memory reads: 64 FMAs per 8 reads; k-split partials are summed in a fixed order (bit-reproducible)
//     layer 3: O <= 16 outputs from the CTA's own units, partial sums exchanged over DSMEM, every CTA sums the four
//              partials in rank order (identical results in all four)
//   head     : tanh-Normal sample + log-prob / deterministic head + TD3 noise, written into the next net's input row
//
// The same kernel serves the acting path (Agent.forward / sample_action at B = num_envs: agent.py:204-327) and any
// forward that does not keep activations.  sm_100a (thread-block clusters, DSMEM, PDL).
#include <cstring>

#include "ssac_mlp.cuh"

namespace ssac {
namespace rows {

constexpr int R = 8;        // batch rows per cluster
constexpr int CS = 4;       // CTAs per cluster = slices of the hidden layers
constexpr int T = 256;      // threads per CTA
constexpr int kMaxH = 256, kMaxO = 16, kMaxD = 64;
constexpr int kXP = kMaxD + 1;   // pitch of the input rows

struct Net {
  const float *W1, *b1, *W2, *b2, *W3, *b3;   // first net of the stack; nets are contiguous per array
  int D, O;
};

struct Args {
  Net actor; int has_actor, det, A, S;
  const float *eps, *noise; float sigma, clip, lo, hi;
  float *logp, *tanh_out, *actor_out;   // nullable outputs: log-prob [B], tanh(out) [B,A] (det), raw head input [B,O]
  float* a_out; int64_t lda;            // sampled action [B,A] (row stride lda), e.g. the action columns of x
  Net critic; const int32_t* net_index; int M;
  float* y;                             // [M,B,O] critic outputs
  const float* x; int64_t ldx; int B, H;
};

#define SSAC_LOG2F 0.6931471805599453f
#define SSAC_LOG_SQRT_2PIF 0.9189385332046727f
__device__ __forceinline__ float softplus_th(float z) { return z > 20.f ? z : log1pf(expf(z)); }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_remote_f4(const float* local_ptr, uint32_t cta, float4 v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr(local_ptr)), "r"(cta));
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_remote_f1(const float* local_ptr, uint32_t cta, float v) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr(local_ptr)), "r"(cta));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

struct Smem {
  float xs[R][kXP];               // input rows of the current net: [s | a]
  float hfull[R][kMaxH];          // layer-1 activations of all units (every CTA's slice, via DSMEM)
  float hloc[R][kMaxH / CS + 1];  // this CTA's layer-1 slice (before the broadcast) / layer-2 slice
  float red[2 * R * T];           // layer-2 k-split partials: [j][r][ks][pair]
  float W1s[kMaxH / CS][kXP];     // this CTA's rows of W1
  float W3s[kMaxO][kMaxH / CS + 1];
  float b1s[kMaxH / CS], b2s[kMaxH / CS], b3s[kMaxO];
  float yp[CS][R][kMaxO];         // output-layer partials of the four CTAs
  float out[R][kMaxO];            // the net's outputs (identical in all four CTAs)
};

// One 3-Linear ReLU MLP on the R rows in sm.xs -> sm.out.  Every thread of all four CTAs calls it.
__device__ void run_net(Smem& sm, const Net& n, int g, int H, int rank, int t) {
  const int D = n.D, O = n.O;
  const int HC = H / CS;                 // units per CTA (H is a multiple of 8: HC is even)
  const int n_lo = rank * HC;
  const int NP = HC / 2;                 // unit pairs
  const int KS = T / NP;                 // k splits (threads beyond NP*KS idle in layer 2)
  const int klen = ((H + KS - 1) / KS + 3) & ~3;   // k per split, multiple of 4, <= 32
  const float* W1 = n.W1 + (int64_t)g * H * D;
  const float* W2 = n.W2 + (int64_t)g * H * H;
  const float* W3 = n.W3 + (int64_t)g * O * H;
  // ---- layer-2 weights of this thread: issued first, consumed after the layer-1 exchange -----------------------------
  const int pair = t % NP, ks = t / NP;
  const bool l2_active = ks < KS;
  const int k0 = ks * klen;
  float4 w[2][8];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float* wr = W2 + (int64_t)(n_lo + 2 * pair + j) * H + k0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool ok = l2_active && 4 * i < klen && k0 + 4 * i < H;
      w[j][i] = ok ? __ldg(reinterpret_cast<const float4*>(wr + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // ---- small operands -> shared memory ---------------------------------------------------------------------------
  for (int i = t; i < HC * D; i += T) {
    const int c = i / D, k = i - c * D;
    sm.W1s[c][k] = __ldg(W1 + (int64_t)(n_lo + c) * D + k);
  }
  for (int i = t; i < O * HC; i += T) {
    const int o = i / HC, c = i - o * HC;
    sm.W3s[o][c] = __ldg(W3 + (int64_t)o * H + n_lo + c);
  }
  if (t < HC) {
    sm.b1s[t] = __ldg(n.b1 + (int64_t)g * H + n_lo + t);
    sm.b2s[t] = __ldg(n.b2 + (int64_t)g * H + n_lo + t);
  }
  if (t < O) sm.b3s[t] = __ldg(n.b3 + (int64_t)g * O + t);
  __syncthreads();
  // ---- layer 1: thread = (unit c, row group) ----------------------------------------------------------------------
  for (int o = t; o < R * HC; o += T) {
    const int c = o % HC, r = o / HC;
    float acc = sm.b1s[c];
    for (int k = 0; k < D; ++k) acc = fmaf(sm.xs[r][k], sm.W1s[c][k], acc);
    sm.hloc[r][c] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  // broadcast the slice into every CTA's hfull (16-byte stores; HC is a multiple of 2: pairs as 8-byte halves are
  // avoided by requiring HC % 4 == 0 on this path, else scalar stores)
  if ((HC & 3) == 0) {
    const int q = HC / 4;
    for (int i = t; i < R * q; i += T) {
      const int r = i / q, c4 = (i - r * q) * 4;
      const float4 v = make_float4(sm.hloc[r][c4], sm.hloc[r][c4 + 1], sm.hloc[r][c4 + 2], sm.hloc[r][c4 + 3]);
#pragma unroll
      for (int d = 0; d < CS; ++d) st_remote_f4(&sm.hfull[r][n_lo + c4], (uint32_t)d, v);
    }
  } else {
    for (int i = t; i < R * HC; i += T) {
      const int r = i / HC, c = i - r * HC;
#pragma unroll
      for (int d = 0; d < CS; ++d) st_remote_f1(&sm.hfull[r][n_lo + c], (uint32_t)d, sm.hloc[r][c]);
    }
  }
  cluster_sync_all();
  // ---- layer 2: thread = (unit pair, k split) ----------------------------------------------------------------------
  if (l2_active) {
    float acc[2][R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[0][r] = acc[1][r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (4 * i < klen && k0 + 4 * i < H) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 h = *reinterpret_cast<const float4*>(&sm.hfull[r][k0 + 4 * i]);
          acc[0][r] = fmaf(h.x, w[0][i].x, acc[0][r]); acc[0][r] = fmaf(h.y, w[0][i].y, acc[0][r]);
          acc[0][r] = fmaf(h.z, w[0][i].z, acc[0][r]); acc[0][r] = fmaf(h.w, w[0][i].w, acc[0][r]);
          acc[1][r] = fmaf(h.x, w[1][i].x, acc[1][r]); acc[1][r] = fmaf(h.y, w[1][i].y, acc[1][r]);
          acc[1][r] = fmaf(h.z, w[1][i].z, acc[1][r]); acc[1][r] = fmaf(h.w, w[1][i].w, acc[1][r]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int r = 0; r < R; ++r) sm.red[((j * R + r) * KS + ks) * NP + pair] = acc[j][r];
  }
  __syncthreads();
  for (int o = t; o < R * HC; o += T) {
    const int c = o % HC, r = o / HC;
    const int j = c & 1, pr = c >> 1;
    float s = sm.b2s[c];
    for (int q = 0; q < KS; ++q) s += sm.red[((j * R + r) * KS + q) * NP + pr];   // fixed order
    sm.hloc[r][c] = fmaxf(s, 0.f);
  }
  __syncthreads();
  // ---- layer 3: partial outputs over this CTA's units, exchanged over DSMEM ------------------------------------------
  for (int o = t; o < R * O; o += T) {
    const int oo = o % O, r = o / O;
    float acc = 0.f;
    for (int c = 0; c < HC; ++c) acc = fmaf(sm.hloc[r][c], sm.W3s[oo][c], acc);
#pragma unroll
    for (int d = 0; d < CS; ++d) st_remote_f1(&sm.yp[rank][r][oo], (uint32_t)d, acc);
  }
  cluster_sync_all();
  for (int o = t; o < R * O; o += T) {
    const int oo = o % O, r = o / O;
    float s = sm.b3s[oo];
#pragma unroll
    for (int d = 0; d < CS; ++d) s += sm.yp[d][r][oo];   // rank order: identical in all four CTAs
    sm.out[r][oo] = s;
  }
  __syncthreads();
}

__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(T, 1) mlp_rows_kernel(const __grid_constant__ Args q) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const int t = threadIdx.x;
  const int rank = (int)cluster_rank();
  const int m0 = ((int)blockIdx.x / CS) * R;
  const int H = q.H, B = q.B;
  const int Din = q.has_actor ? q.S : q.critic.D;   // columns of x that are inputs (the actor reads the state columns)
  // the batch (x, eps, noise) is the previous kernels' business: everything else above was parameter-free set-up
  pdl_wait();
  pdl_trigger();
  for (int i = t; i < R * kXP; i += T) (&sm.xs[0][0])[i] = 0.f;
  __syncthreads();
  const int Dload = q.has_actor ? ((q.M > 0) ? q.S : q.actor.D) : Din;
  for (int i = t; i < R * Dload; i += T) {
    const int r = i / Dload, k = i - r * Dload;
    if (m0 + r < B) sm.xs[r][k] = q.x[(int64_t)(m0 + r) * q.ldx + k];
  }
  __syncthreads();

  if (q.has_actor) {
    run_net(sm, q.actor, 0, H, rank, t);
    // ---- policy head: thread = row (all four CTAs compute it, rank 0 writes the global outputs) -----------------------
    if (t < R) {
      const int r = t, b = m0 + r;
      const bool wr = rank == 0 && b < B;
      const int A = q.A;
      if (q.actor_out && wr)
        for (int o = 0; o < q.actor.O; ++o) q.actor_out[(int64_t)b * q.actor.O + o] = sm.out[r][o];
      if (!q.det) {
        // a = tanh(mu + eps*std), logp = sum_j [Normal.log_prob(x_j) - log|d tanh|]  (nets/distributions.py:9-15,64-104)
        float lp = 0.f;
        for (int j = 0; j < A; ++j) {
          const float mu = sm.out[r][j], raw = sm.out[r][A + j];
          const float e = b < B ? q.eps[(int64_t)b * A + j] : 0.f;
          const float t_raw = tanhf(raw);
          const float log_std = q.lo + 0.5f * (q.hi - q.lo) * (t_raw + 1.f);
          const float sd = expf(log_std);
          const float xv = mu + e * sd;
          const float av = tanhf(xv);
          const float ladj = 2.f * (SSAC_LOG2F - xv - softplus_th(-2.f * xv));
          const float dxm = xv - mu;
          lp += (0.f - ladj) + (-(dxm * dxm) / (2.f * (sd * sd)) - logf(sd) - SSAC_LOG_SQRT_2PIF);
          sm.xs[r][q.S + j] = av;
          if (wr && q.a_out) q.a_out[(int64_t)b * q.lda + j] = av;
        }
        if (wr && q.logp) q.logp[b] = lp;
      } else {
        // deterministic head (+ rsample jitter, + TD3 noise with the straight-through clamp: learning_utils.py:48-59)
        for (int j = 0; j < A; ++j) {
          const int64_t i = (int64_t)b * A + j;
          const float th = tanhf(sm.out[r][j]);
          if (wr && q.tanh_out) q.tanh_out[i] = th;
          float v2 = th;
          if (q.eps && b < B) v2 = __fadd_rn(v2, __fmul_rn(q.eps[i], 1e-4f));
          if (q.noise && b < B) {
            float nz = __fmul_rn(q.sigma, q.noise[i]);
            if (q.clip > 0.f) nz = fminf(fmaxf(nz, -q.clip), q.clip);
            v2 = __fadd_rn(v2, nz);
            v2 = fminf(fmaxf(v2, __fadd_rn(-1.f, 1e-6f)), __fadd_rn(1.f, -1e-6f));
          }
          sm.xs[r][q.S + j] = v2;
          if (wr && q.a_out) q.a_out[(int64_t)b * q.lda + j] = v2;
        }
      }
    }
    __syncthreads();
  }
  for (int m = 0; m < q.M; ++m) {
    const int g = q.net_index ? q.net_index[m] : m;
    run_net(sm, q.critic, g, H, rank, t);
    const int O = q.critic.O;
    if (rank == 0)
      for (int o = t; o < R * O; o += T) {
        const int oo = o % O, r = o / O;
        if (m0 + r < B) q.y[((int64_t)m * B + m0 + r) * O + oo] = sm.out[r][oo];
      }
    __syncthreads();
  }
  cluster_sync_all();   // nobody leaves while a neighbour may still write into its shared memory
}

static int launch(const Args& a, cudaStream_t s) {
  static bool attr_set = false;
  const size_t smem = sizeof(Smem);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error(std::string("mlp_rows (smem attribute): ") + cudaGetErrorString(e)); return (int)e; }
    attr_set = true;
  }
  const int blocks = (a.B + R - 1) / R;
  cudaError_t le = launch_pdl(mlp_rows_kernel, dim3(CS * blocks), dim3(T), smem, s, a);
  if (le != cudaSuccess) { set_error(std::string("mlp_rows: ") + cudaGetErrorString(le)); return (int)le; }
  SSAC_CHECK_LAUNCH("mlp_rows");
  return 0;
}

bool shape_ok(int D, int H, int O) {
  return H >= 32 && H <= kMaxH && (H % 8) == 0 && D >= 1 && D <= kMaxD && O >= 1 && O <= kMaxO;
}

}  // namespace rows

// Plain forward of G nets on a shared batch without kept activations (impl 3 of ssac_mlp_forward).
int mlp_forward_rows(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3, const float* b3,
                     const int32_t* net_index, int G, int D, int H, int O, const float* x, int64_t ldx, int B, float* y,
                     cudaStream_t s) {
  if (!rows::shape_ok(D, H, O)) return fail(SSAC_E_UNSUPPORTED, "mlp rows kernel: H in 32..256 (multiple of 8), D <= 64, O <= 16");
  rows::Args a;
  memset(&a, 0, sizeof(a));
  a.critic = rows::Net{W1, b1, W2, b2, W3, b3, D, O};
  a.net_index = net_index; a.M = G; a.y = y; a.x = x; a.ldx = ldx; a.B = B; a.H = H;
  return rows::launch(a, s);
}

}  // namespace ssac

using namespace ssac;

extern "C" {

int ssac_policy_rows(const float* W1, const float* b1, const float* W2, const float* b2, const float* W3, const float* b3,
                     int D, int H, int A, int deterministic, const float* x_dev, int64_t ldx, int B, float* out_dev,
                     const float* eps_dev, const float* noise_dev, float sigma, float clip, float log_std_lo,
                     float log_std_hi, float* a_dev, int64_t lda, float* logp_dev, float* tanh_out_dev, void* stream) {
  SSAC_REQUIRE(W1 && b1 && W2 && b2 && W3 && b3 && x_dev, "ssac_policy_rows: null pointer");
  SSAC_REQUIRE(D > 0 && A > 0 && A <= 8 && B > 0 && ldx >= D, "ssac_policy_rows: bad sizes");
  SSAC_REQUIRE(deterministic || eps_dev, "ssac_policy_rows: a stochastic actor needs eps");
  const int O = deterministic ? A : 2 * A;
  if (!rows::shape_ok(D, H, O) || D + A > rows::kMaxD) return fail(SSAC_E_UNSUPPORTED, "ssac_policy_rows: shape outside the rows kernel");
  rows::Args a;
  memset(&a, 0, sizeof(a));
  a.actor = rows::Net{W1, b1, W2, b2, W3, b3, D, O};
  a.has_actor = 1; a.det = deterministic; a.A = A; a.S = D;
  a.eps = eps_dev; a.noise = noise_dev; a.sigma = sigma; a.clip = clip; a.lo = log_std_lo; a.hi = log_std_hi;
  a.logp = logp_dev; a.tanh_out = tanh_out_dev; a.actor_out = out_dev; a.a_out = a_dev; a.lda = lda;
  a.M = 0; a.x = x_dev; a.ldx = ldx; a.B = B; a.H = H;
  return rows::launch(a, (cudaStream_t)stream);
}

int ssac_target_chain(const float* aW1, const float* ab1, const float* aW2, const float* ab2, const float* aW3,
                      const float* ab3, int S, int H, int A, int deterministic, const float* cW1, const float* cb1,
                      const float* cW2, const float* cb2, const float* cW3, const float* cb3,
                      const int32_t* net_index_dev, int M, float* x1_dev, int64_t ldx, int B, const float* eps_dev,
                      const float* noise_dev, float sigma, float clip, float log_std_lo, float log_std_hi,
                      float* logp_dev, float* qt_dev, void* stream) {
  SSAC_REQUIRE(aW1 && ab1 && aW2 && ab2 && aW3 && ab3 && cW1 && cb1 && cW2 && cb2 && cW3 && cb3 && x1_dev && qt_dev,
               "ssac_target_chain: null pointer");
  SSAC_REQUIRE(S > 0 && A > 0 && A <= 8 && M > 0 && B > 0 && ldx >= S + A, "ssac_target_chain: bad sizes");
  SSAC_REQUIRE(deterministic || (eps_dev && logp_dev), "ssac_target_chain: a stochastic actor needs eps and logp");
  const int O = deterministic ? A : 2 * A;
  if (!rows::shape_ok(S + A, H, O)) return fail(SSAC_E_UNSUPPORTED, "ssac_target_chain: shape outside the rows kernel");
  rows::Args a;
  memset(&a, 0, sizeof(a));
  a.actor = rows::Net{aW1, ab1, aW2, ab2, aW3, ab3, S, O};
  a.has_actor = 1; a.det = deterministic; a.A = A; a.S = S;
  a.eps = eps_dev; a.noise = noise_dev; a.sigma = sigma; a.clip = clip; a.lo = log_std_lo; a.hi = log_std_hi;
  a.logp = logp_dev; a.a_out = x1_dev + S; a.lda = ldx;
  a.critic = rows::Net{cW1, cb1, cW2, cb2, cW3, cb3, S + A, 1};
  a.net_index = net_index_dev; a.M = M; a.y = qt_dev; a.x = x1_dev; a.ldx = ldx; a.B = B; a.H = H;
  return rows::launch(a, (cudaStream_t)stream);
}

static int g_rows_enabled = 1;
int ssac_rows_supported(int D, int H, int O) { return (g_rows_enabled && rows::shape_ok(D, H, O)) ? 1 : 0; }
int ssac_set_rows_enabled(int on) { g_rows_enabled = on ? 1 : 0; return 0; }

}  // extern "C"
