// SAC-Discrete heads: the categorical-policy arithmetic of the reference's discrete branches
// (learning_utils.py:322-328, learning.py:84-92, :252-253, :382-390) as row-local kernels.  The networks themselves are
// the same grouped MLP launches as the continuous path (O = number of actions); what is left is per batch row a softmax
// over A logits, a min over the critics' Q rows and an expectation -- A is small (Atari: <= 18), so one warp owns one row,
// lanes stride over the actions, and everything stays in registers.  All of it is a few KB per launch: latency-bound glue,
// kept off ATen so the update is the repo's own launches end to end.
#include "ssac_common.cuh"

namespace ssac {
namespace {

constexpr int kRowsPerBlock = 4;   // warps per block, one batch row each

struct RowSoftmax {
  float zmax, lse;   // log p_a = z_a - zmax - lse
};
// log-softmax statistics of one row, lanes striding over the A logits (every lane gets the result)
__device__ __forceinline__ RowSoftmax row_softmax(const float* __restrict__ z, int A, int lane) {
  float m = -INFINITY;
  for (int a = lane; a < A; a += 32) m = fmaxf(m, z[a]);
  m = warp_max(m);
  float s = 0.f;
  for (int a = lane; a < A; a += 32) s += expf(z[a] - m);
  s = warp_sum(s);
  return RowSoftmax{m, logf(s)};
}

// learning_utils.py:322-328: v[b] = sum_a p_a (min_m Q_m[b,a] - alpha log p_a);  ent += sum_{b,a} alpha log p_a / (B A)
__global__ void __launch_bounds__(32 * kRowsPerBlock) discrete_value_kernel(const float* __restrict__ logits,
                                                                           const float* __restrict__ qt, int M, int B,
                                                                           int A, const float* __restrict__ log_alpha,
                                                                           float* __restrict__ v,
                                                                           float* __restrict__ ent) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (b >= B) return;
  const float alpha = expf(*log_alpha);
  const float* z = logits + (int64_t)b * A;
  const RowSoftmax sm = row_softmax(z, A, lane);
  float acc = 0.f, e = 0.f;
  for (int a = lane; a < A; a += 32) {
    const float lp = z[a] - sm.zmax - sm.lse, p = expf(lp);
    float q = qt[(int64_t)b * A + a];
    for (int m = 1; m < M; ++m) q = fminf(q, qt[((int64_t)m * B + b) * A + a]);
    const float bonus = alpha * lp;
    acc += p * (q - bonus);
    e += bonus;
  }
  acc = warp_sum(acc);
  e = warp_sum(e);
  if (lane == 0) {
    v[b] = acc;
    if (ent) atomicAdd(ent, e / ((float)B * (float)A));
  }
}

// out[g,b] = q[g,b,a_b]  (q.gather(-1, a.long()): learning.py:91, learning_utils.py:373-376)
__global__ void discrete_gather_q_kernel(const float* __restrict__ q, const float* __restrict__ act, int G, int B, int A,
                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * B) return;
  const int b = i % B;
  int a = (int)act[b];
  a = a < 0 ? 0 : (a >= A ? A - 1 : a);
  out[i] = q[(int64_t)i * A + a];
}

// learning.py:84-98, :112 with discrete=True.  One block (B*N*A is a few thousand elements), as critic_loss_seed_kernel:
// q' = popart(q[k,b,a_b]); dy[k,b,a] = -2 w imp (y - q') popw / (B E n_total) at a = a_b, 0 elsewhere;
// loss[0] += sum_k mean_b(w imp (y-q')^2) / (E n_total); loss[1] = mean_b(y - q'_{N-1}).
__global__ void __launch_bounds__(1024) discrete_critic_loss_seed_kernel(const float* __restrict__ q, int N, int B, int A,
                                                                        const float* __restrict__ act,
                                                                        const float* __restrict__ y,
                                                                        const float* __restrict__ w,
                                                                        const float* __restrict__ imp,
                                                                        const float* __restrict__ popart, int pop, int E,
                                                                        int n_total, float* __restrict__ dy,
                                                                        float* __restrict__ loss) {
  __shared__ float scratch[32];
  const float pw = (popart && pop) ? popart[2] : 1.f, pb = (popart && pop) ? popart[3] : 0.f;
  const float inv_count = 1.f / ((float)B * (float)E * (float)n_total);
  float s_loss = 0.f, s_td = 0.f;
  for (int i = threadIdx.x; i < N * B; i += blockDim.x) {
    const int k = i / B, b = i - k * B;
    int ab = (int)act[b];
    ab = ab < 0 ? 0 : (ab >= A ? A - 1 : ab);
    const float qs = q[(int64_t)i * A + ab];
    const float qq = (popart && pop) ? __fadd_rn(__fmul_rn(pw, qs), pb) : qs;
    const float td = y[b] - qq;
    const float ww = (w ? w[b] : 1.f) * (imp ? imp[b] : 1.f);
    s_loss += ww * td * td;
    if (k == N - 1) s_td += td;
    const float seed = -2.f * ww * td * pw * inv_count;
    for (int a = 0; a < A; ++a) dy[(int64_t)i * A + a] = (a == ab) ? seed : 0.f;
  }
  s_loss = block_reduce(s_loss, scratch, OpSum(), 0.f);
  s_td = block_reduce(s_td, scratch, OpSum(), 0.f);
  if (threadIdx.x == 0 && loss) {
    atomicAdd(&loss[0], s_loss * inv_count);
    loss[1] = s_td / (float)B;
  }
}

// learning.py:382-390, :408-409: vals_a = popart(min_n Q_n[b,a]); g_a = vals_a - alpha log p_a; f = sum_a p_a g_a;
// loss += -(1/E) mean_b f;  d loss / d logit_k = -(1 / (E B)) p_k (g_k - f)   (the alpha * sum_a p_a dlogp_a term is 0).
__global__ void __launch_bounds__(32 * kRowsPerBlock) discrete_actor_seed_kernel(const float* __restrict__ logits,
                                                                                const float* __restrict__ q, int N,
                                                                                int B, int A,
                                                                                const float* __restrict__ log_alpha,
                                                                                const float* __restrict__ popart, int pop,
                                                                                int E, float* __restrict__ dlogits,
                                                                                float* __restrict__ loss) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (b >= B) return;
  const float alpha = expf(*log_alpha);
  const float pw = (popart && pop) ? popart[2] : 1.f, pb = (popart && pop) ? popart[3] : 0.f;
  const float* z = logits + (int64_t)b * A;
  const RowSoftmax sm = row_softmax(z, A, lane);
  float f = 0.f;
  for (int a = lane; a < A; a += 32) {
    const float lp = z[a] - sm.zmax - sm.lse, p = expf(lp);
    float qq = q[(int64_t)b * A + a];
    for (int n = 1; n < N; ++n) qq = fminf(qq, q[((int64_t)n * B + b) * A + a]);
    if (popart && pop) qq = __fadd_rn(__fmul_rn(pw, qq), pb);
    f += p * (qq - alpha * lp);
  }
  f = warp_sum(f);
  const float scale = -1.f / ((float)E * (float)B);
  for (int a = lane; a < A; a += 32) {
    const float lp = z[a] - sm.zmax - sm.lse, p = expf(lp);
    float qq = q[(int64_t)b * A + a];
    for (int n = 1; n < N; ++n) qq = fminf(qq, q[((int64_t)n * B + b) * A + a]);
    if (popart && pop) qq = __fadd_rn(__fmul_rn(pw, qq), pb);
    dlogits[(int64_t)b * A + a] = scale * p * ((qq - alpha * lp) - f);
  }
  if (lane == 0 && loss) atomicAdd(loss, scale * f);
}

// learning.py:252-253: out[b] = sum_a p_a log p_a
__global__ void __launch_bounds__(32 * kRowsPerBlock) discrete_neg_entropy_kernel(const float* __restrict__ logits, int B,
                                                                                 int A, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* z = logits + (int64_t)b * A;
  const RowSoftmax sm = row_softmax(z, A, lane);
  float s = 0.f;
  for (int a = lane; a < A; a += 32) {
    const float lp = z[a] - sm.zmax - sm.lse;
    s += expf(lp) * lp;
  }
  s = warp_sum(s);
  if (lane == 0) out[b] = s;
}

// adv_estimator.py:45-56 (discrete 'indirect') + learning_utils.py:257-262, :288-295.  logits [E,B,A]: EVERY actor of the
// ensemble on the batch (V uses the ensemble-mean policy), q [N,B,A]: the member's critics.  min_q = popart(min_N q) when the
// member has a PopArt layer (adv_estimator.py:30-35 has no `pop` switch); V = sum_a mean_e p_e,a min_q_a;
// adv = min_q[a_b] - V; mask = adv >= 0; priority (float64) = relu(adv) + 1e-4.
__global__ void __launch_bounds__(32 * kRowsPerBlock) discrete_advantage_kernel(const float* __restrict__ logits, int E,
                                                                               const float* __restrict__ q, int N, int B,
                                                                               int A, const float* __restrict__ act,
                                                                               const float* __restrict__ popart,
                                                                               float* __restrict__ adv,
                                                                               float* __restrict__ mask,
                                                                               double* __restrict__ prio) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (b >= B) return;
  const float pw = popart ? popart[2] : 1.f, pb = popart ? popart[3] : 0.f;
  int ab = (int)act[b];
  ab = ab < 0 ? 0 : (ab >= A ? A - 1 : ab);
  float value = 0.f, q_data = 0.f;
  for (int e = 0; e < E; ++e) {
    const float* z = logits + ((int64_t)e * B + b) * A;
    const RowSoftmax sm = row_softmax(z, A, lane);
    for (int a = lane; a < A; a += 32) {
      float qq = q[(int64_t)b * A + a];
      for (int n = 1; n < N; ++n) qq = fminf(qq, q[((int64_t)n * B + b) * A + a]);
      if (popart) qq = __fadd_rn(__fmul_rn(pw, qq), pb);
      value += expf(z[a] - sm.zmax - sm.lse) * qq;
      if (e == 0 && a == ab) q_data = qq;
    }
  }
  value = warp_sum(value) / (float)E;
  q_data = warp_sum(q_data);   // exactly one lane holds it
  if (lane == 0) {
    const float ad = q_data - value;
    if (adv) adv[b] = ad;
    if (mask) mask[b] = ad >= 0.f ? 1.f : 0.f;
    if (prio) prio[b] = (double)fmaxf(ad, 0.f) + 1e-4;
  }
}

// learning_utils.py:257-269 with discrete=True: logp = log p_{a_b}; loss[0] += -(1/B) sum_b mask_b logp_b (the member's
// filtered BC loss; mask nullable = 1); dlogits[b,k] = -(mask_b / (B E)) (1[k = a_b] - p_k).
__global__ void __launch_bounds__(32 * kRowsPerBlock) discrete_bc_seed_kernel(const float* __restrict__ logits,
                                                                             const float* __restrict__ act,
                                                                             const float* __restrict__ mask, int B, int A,
                                                                             int E, float* __restrict__ dlogits,
                                                                             float* __restrict__ loss) {
  const int lane = threadIdx.x & 31, b = blockIdx.x * kRowsPerBlock + (threadIdx.x >> 5);
  if (b >= B) return;
  int ab = (int)act[b];
  ab = ab < 0 ? 0 : (ab >= A ? A - 1 : ab);
  const float m = mask ? mask[b] : 1.f;
  const float* z = logits + (int64_t)b * A;
  const RowSoftmax sm = row_softmax(z, A, lane);
  const float scale = -m / ((float)B * (float)E);
  for (int a = lane; a < A; a += 32) {
    const float p = expf(z[a] - sm.zmax - sm.lse);
    dlogits[(int64_t)b * A + a] = scale * ((a == ab ? 1.f : 0.f) - p);
  }
  if (lane == 0 && loss) atomicAdd(loss, -m * (z[ab] - sm.zmax - sm.lse) / (float)B);
}

inline int row_grid(int B) { return (B + kRowsPerBlock - 1) / kRowsPerBlock; }

}  // namespace
}  // namespace ssac

using namespace ssac;

extern "C" {

int ssac_discrete_value(const float* logits, const float* q_t, int M, int B, int A, const float* log_alpha, float* v,
                        float* ent, void* stream) {
  SSAC_REQUIRE(logits && q_t && log_alpha && v && M > 0 && B > 0 && A > 0, "ssac_discrete_value: bad args");
  discrete_value_kernel<<<row_grid(B), 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(logits, q_t, M, B, A, log_alpha, v,
                                                                                     ent);
  SSAC_CHECK_LAUNCH("ssac_discrete_value");
  return 0;
}

int ssac_discrete_gather_q(const float* q, const float* act, int G, int B, int A, float* out, void* stream) {
  SSAC_REQUIRE(q && act && out && G > 0 && B > 0 && A > 0, "ssac_discrete_gather_q: bad args");
  discrete_gather_q_kernel<<<(G * B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(q, act, G, B, A, out);
  SSAC_CHECK_LAUNCH("ssac_discrete_gather_q");
  return 0;
}

int ssac_discrete_critic_loss_seed(const float* q, int N, int B, int A, const float* act, const float* y, const float* w,
                                   const float* imp, const float* popart, int pop, int E, int n_total, float* dy,
                                   float* loss, void* stream) {
  SSAC_REQUIRE(q && act && y && dy && N > 0 && B > 0 && A > 0 && E > 0 && n_total >= 0,
               "ssac_discrete_critic_loss_seed: bad args");
  discrete_critic_loss_seed_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(q, N, B, A, act, y, w, imp, popart, pop, E,
                                                                         n_total > 0 ? n_total : N, dy, loss);
  SSAC_CHECK_LAUNCH("ssac_discrete_critic_loss_seed");
  return 0;
}

int ssac_discrete_actor_seed(const float* logits, const float* q, int N, int B, int A, const float* log_alpha,
                             const float* popart, int pop, int E, float* dlogits, float* loss, void* stream) {
  SSAC_REQUIRE(logits && q && log_alpha && dlogits && N > 0 && B > 0 && A > 0 && E > 0,
               "ssac_discrete_actor_seed: bad args");
  discrete_actor_seed_kernel<<<row_grid(B), 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(logits, q, N, B, A, log_alpha,
                                                                                          popart, pop, E, dlogits, loss);
  SSAC_CHECK_LAUNCH("ssac_discrete_actor_seed");
  return 0;
}

int ssac_discrete_neg_entropy(const float* logits, int B, int A, float* out, void* stream) {
  SSAC_REQUIRE(logits && out && B > 0 && A > 0, "ssac_discrete_neg_entropy: bad args");
  discrete_neg_entropy_kernel<<<row_grid(B), 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(logits, B, A, out);
  SSAC_CHECK_LAUNCH("ssac_discrete_neg_entropy");
  return 0;
}

int ssac_discrete_advantage(const float* logits, int E, const float* q, int N, int B, int A, const float* act,
                            const float* popart, float* adv, float* mask, double* priority, void* stream) {
  SSAC_REQUIRE(logits && q && act && E > 0 && N > 0 && B > 0 && A > 0, "ssac_discrete_advantage: bad args");
  discrete_advantage_kernel<<<row_grid(B), 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(logits, E, q, N, B, A, act,
                                                                                         popart, adv, mask, priority);
  SSAC_CHECK_LAUNCH("ssac_discrete_advantage");
  return 0;
}

int ssac_discrete_bc_seed(const float* logits, const float* act, const float* mask, int B, int A, int E, float* dlogits,
                          float* loss, void* stream) {
  SSAC_REQUIRE(logits && act && dlogits && B > 0 && A > 0 && E > 0, "ssac_discrete_bc_seed: bad args");
  discrete_bc_seed_kernel<<<row_grid(B), 32 * kRowsPerBlock, 0, (cudaStream_t)stream>>>(logits, act, mask, B, A, E, dlogits,
                                                                                       loss);
  SSAC_CHECK_LAUNCH("ssac_discrete_bc_seed");
  return 0;
}

}  // extern "C"
