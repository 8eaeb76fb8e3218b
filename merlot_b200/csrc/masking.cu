// K12: attention-guided MLM mask selection (model/modeling.py:381-489) as one block per sequence, entirely on device.
// Integer/indexing path: bit-exact against the oracle given the same attention sums and the same injected random draws
// (Gumbel noise, two SpanBERT span draws, 10/80/10 option draw, random replacement ids).
//
// tf.math.top_k semantics (ties -> lower index first) are realised by rank counting:
//   rank(l) = #{l' : x[l'] > x[l]  or  (x[l'] == x[l] and l' < l)},   member of top-k <=> rank < k.
// Every floating-point expression uses explicit round-to-nearest mul/add/div intrinsics (no FMA contraction) so that it
// rounds exactly like the reference's op-by-op fp32 graph.
#include "host_common.h"
#include "ptx.cuh"

namespace mb {

struct MaskDev {
  const int* ids;            // [B, L]
  const float* attn_summ;    // [B, L] or null (masking_use_attn = False)
  const float* gumbel;       // [B, L]
  const int* span_lower;     // [B, k]
  const int* span_upper;     // [B, k]
  const int* option;         // [B*L] in {0,1,2}
  const int* rand_ids;       // [B*L]
  int* masked_ids;           // [B, L]
  int* masked_idx;           // [B, k] ascending
  uint8_t* valid_out;        // [B, L] masked_ids != 0 (optional)
  int B, L, num_topk, num_to_mask, do_spanbert, mask_token;
  float w_delta, w_non, logw_top, logw_non, w_max;
};

__global__ void __launch_bounds__(1024) mask_inputs_kernel(const MaskDev p) {
  extern __shared__ float sm[];
  const int L = p.L, k = p.num_to_mask;
  float* val = sm;                 // [L] scratch values to rank
  float* w = sm + L;               // [L] mask_weight
  int* idx = reinterpret_cast<int*>(sm + 2 * L);  // [k] Gumbel-top-k indices (reversed order)
  int* member = idx + k;           // [L] final membership
  const int b = blockIdx.x;
  const int* ids = p.ids + (size_t)b * L;

  // ---- 1-3: importance top-k -> mask_weight ----
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float special = ids[l] < 100 ? 1.f : 0.f;
    val[l] = p.attn_summ ? __fmul_rn(p.attn_summ[(size_t)b * L + l], __fsub_rn(1.f, special)) : 0.f;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float wl = 1.f;
    if (p.attn_summ) {
      const float x = val[l];
      int rank = 0;
      for (int j = 0; j < L; ++j) { const float y = val[j]; rank += (y > x) || (y == x && j < l); }
      wl = __fadd_rn(__fmul_rn(rank < p.num_topk ? 1.f : 0.f, p.w_delta), p.w_non);
    }
    w[l] = wl;
  }
  __syncthreads();
  // ---- 4-5: Gumbel top-k without replacement, reversed ----
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float special = ids[l] < 100 ? 1.f : 0.f;
    float lw = 0.f;
    if (p.attn_summ) lw = (w[l] == __fadd_rn(p.w_delta, p.w_non)) ? p.logw_top : p.logw_non;
    const float log_mask = __fsub_rn(lw, __fmul_rn(1e8f, special));
    val[l] = __fadd_rn(log_mask, p.gumbel[(size_t)b * L + l]);
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float x = val[l];
    int rank = 0;
    for (int j = 0; j < L; ++j) { const float y = val[j]; rank += (y > x) || (y == x && j < l); }
    if (rank < k) idx[k - 1 - rank] = l;  // [:, ::-1]
  }
  __syncthreads();
  // ---- 6: SpanBERT expansion ----
  if (p.do_spanbert) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float special = ids[l] < 100 ? 1.f : 0.f;
      int which = 0;
      for (int m = 0; m < k; ++m) {
        const int s0 = idx[m] - p.span_lower[(size_t)b * k + m], s1 = idx[m] + p.span_upper[(size_t)b * k + m];
        if (l >= s0 && l <= s1) { which = m; break; }  // argmax of a 0/1 vector: first match, 0 if none
      }
      const float wm = __fmul_rn((float)which, __fsub_rn(1.f, special));
      val[l] = __fadd_rn(wm, __fdiv_rn(__fmul_rn(0.5f, w[l]), p.w_max));
    }
    __syncthreads();
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const float x = val[l];
      int rank = 0;
      for (int j = 0; j < L; ++j) { const float y = val[j]; rank += (y > x) || (y == x && j < l); }
      member[l] = rank < k;
    }
  } else {
    for (int l = threadIdx.x; l < L; l += blockDim.x) member[l] = 0;
    __syncthreads();
    for (int m = threadIdx.x; m < k; m += blockDim.x) member[idx[m]] = 1;
  }
  __syncthreads();
  // ---- 7-8: sorted mask_idx, 10/80/10 replacement ----
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    int out = ids[l];
    if (member[l]) {
      int pos = 0;
      for (int j = 0; j < l; ++j) pos += member[j];
      p.masked_idx[(size_t)b * k + pos] = l;
      const int opt = p.option[(size_t)b * L + l];
      out = opt == 0 ? ids[l] : (opt == 1 ? p.mask_token : p.rand_ids[(size_t)b * L + l]);
    }
    p.masked_ids[(size_t)b * L + l] = out;
    if (p.valid_out) p.valid_out[(size_t)b * L + l] = out != 0;
  }
}

__global__ void ids_valid_kernel(const int* __restrict__ ids, uint8_t* __restrict__ valid, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) valid[i] = ids[i] != 0;
}

// The five tf.random draws of mask_inputs (model/modeling.py:445-481, utils/model_utils.py:640-649) on device: counter-based
// Philox keyed by (seed, tensor id), so a step is reproducible from its seed and no host RNG / pageable host->device copy sits
// in the step (the reference draws them inside the TF graph as well).  Same distributions as the reference; the bit streams of
// tf.random cannot be reproduced.
//   gumbel  [B*L]  = -log(-log(U)), U uniform in [1e-9, 1 - 1e-7]      span_lower/upper [B*k] ~ categorical(span_probs)
//   option  [B*L]  ~ categorical(0.1 keep, 0.8 [MASK], 0.1 random)     rand_ids [B*L] uniform in [100, vocab)
__global__ void mask_draws_kernel(float* __restrict__ gumbel, int* __restrict__ span_lower, int* __restrict__ span_upper,
                                  int* __restrict__ option, int* __restrict__ rand_ids, long long n_tok, long long n_span, int vocab,
                                  float p0, float p1, uint64_t seed) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  auto u01 = [](uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); };  // (0,1), 24 bits
  if (i < n_tok) {
    const uint4 r = philox4x32_7(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0x6d61736bu, 1u), key);
    const float u = fminf(fmaxf(u01(r.x), 1e-9f), 1.0f - 1e-7f);
    gumbel[i] = -logf(-logf(u));
    const float o = u01(r.y);
    option[i] = o < 0.1f ? 0 : (o < 0.9f ? 1 : 2);
    rand_ids[i] = 100 + (int)(((uint64_t)r.z * (uint64_t)(vocab - 100)) >> 32);
  }
  if (i < n_span) {
    const uint4 r = philox4x32_7(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0x7370616eu, 2u), key);
    const float a = u01(r.x), b = u01(r.y);
    span_lower[i] = a < p0 ? 0 : (a < p0 + p1 ? 1 : 2);
    span_upper[i] = b < p0 ? 0 : (b < p0 + p1 ? 1 : 2);
  }
}

}  // namespace mb

using namespace mb;

extern "C" int merlot_mask_draws(float* gumbel, int* span_lower, int* span_upper, int* option, int* rand_ids, long long n_tok,
                                 long long n_span, int vocab, float p_len0, float p_len1, uint64_t seed, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(gumbel && span_lower && span_upper && option && rand_ids, MERLOT_EINVAL, "mask_draws: null output");
  MB_REQUIRE(n_tok > 0 && n_span >= 0 && vocab > 100, MERLOT_ESHAPE, "mask_draws: bad sizes (vocab must exceed 100)");
  const long long n = n_tok > n_span ? n_tok : n_span;
  mask_draws_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(gumbel, span_lower, span_upper, option, rand_ids, n_tok, n_span, vocab,
                                                                     p_len0, p_len1, seed);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

namespace mb {
}  // namespace mb

using namespace mb;

extern "C" int merlot_mask_inputs(const merlot_mask_t* m, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(m && m->ids && m->gumbel && m->option && m->rand_ids && m->masked_ids && m->masked_idx, MERLOT_EINVAL,
             "mask_inputs: null pointer");
  MB_REQUIRE(!m->do_spanbert || (m->span_lower && m->span_upper), MERLOT_EINVAL, "mask_inputs: span draws required");
  MB_REQUIRE(m->L > 0 && m->num_to_mask > 0 && m->num_to_mask <= m->L && m->num_topk <= m->L, MERLOT_ESHAPE,
             "mask_inputs: need 0 < num_to_mask <= L (L=%d, num_to_mask=%d, num_topk=%d)", m->L, m->num_to_mask, m->num_topk);
  MaskDev p;
  p.ids = m->ids; p.attn_summ = m->attn_summ; p.gumbel = m->gumbel; p.span_lower = m->span_lower; p.span_upper = m->span_upper;
  p.option = m->option; p.rand_ids = m->rand_ids; p.masked_ids = m->masked_ids; p.masked_idx = m->masked_idx;
  p.valid_out = reinterpret_cast<uint8_t*>(m->valid_out);
  p.B = m->B; p.L = m->L; p.num_topk = m->num_topk; p.num_to_mask = m->num_to_mask; p.do_spanbert = m->do_spanbert;
  p.mask_token = m->mask_token;
  p.w_delta = m->w_delta; p.w_non = m->w_non; p.logw_top = m->logw_top; p.logw_non = m->logw_non; p.w_max = m->w_max;
  const size_t smem = (size_t)(3 * m->L + m->num_to_mask) * 4;
  MB_REQUIRE(smem <= 200 * 1024, MERLOT_ESHAPE, "mask_inputs: sequence too long for one block (L=%d)", m->L);
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(mask_inputs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int threads = m->L >= 1024 ? 1024 : ((m->L + 31) / 32) * 32;
  mask_inputs_kernel<<<m->B, threads, smem, st>>>(p);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

extern "C" int merlot_ids_valid(const int* ids, void* valid, long long n, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(ids && valid, MERLOT_EINVAL, "ids_valid: null pointer");
  if (n == 0) return MERLOT_OK;
  ids_valid_kernel<<<(unsigned)ceil_div_ll(n, 256), 256, 0, st>>>(ids, reinterpret_cast<uint8_t*>(valid), n);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}
