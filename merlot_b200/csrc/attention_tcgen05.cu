// K2/K3/K4: masked softmax attention on tcgen05 tensor cores, FlashAttention-style (probabilities never hit HBM).
//
// Replaces utils/transformer.py:98-127 (scores = q k^T / sqrt(d); scores*m - 1e10*(1-m); softmax; probs @ v) and its
// tf.gradients, plus the consumers of the materialised probabilities: the head-mean column sums that
// model/modeling.py:428 (mask_inputs) takes from `self_attn_probs` (utils/transformer.py:208-209,238).
//
// Layout: q/k/v are read in place from the fused QKV GEMM output [tokens, 3H] (columns [0,H) = q, [H,2H) = k,
// [2H,3H) = v, head h at column h*64) through one 2-D TMA map; ctx / d_ctx are [tokens, H].  Head size is 64.
//
// Mask semantics (reference :109-112): m[q,k] = valid[q] & valid[k].  A masked entry's score is
// exactly -1e10; a padding QUERY row therefore has all scores equal and softmaxes to uniform 1/S over all S keys.
// We realise that row as all-zero scores (identical softmax, but keeps log-sum-exp = log S representable).
#include "host_common.h"
#include "ptx.cuh"

namespace mb {

constexpr int AT_M = 128;   // query rows per tile (TMEM lanes)
constexpr int AT_N = 128;   // keys per tile
constexpr int AT_D = 64;    // head size
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

struct AttnDev {
  int B, S, heads, H;
  const uint8_t* valid;  // [B*S] or null (all valid)
  float scale;
  bf16* ctx; int ld_ctx;         // fwd out [B*S, H]
  float* lse;                    // [B, heads, S] natural-log LSE of the masked, scaled scores
  // backward
  float* dsum;                   // [B, heads, S]  D = rowsum(dO * O)
  float* dq_accum; int ld_dq;    // fp32 [parts][B*S, H]: per-key-tile slices (or one atomically accumulated slice)
  size_t dq_part_stride;         // elements between slices
  bf16* dqkv; int ld_dqkv;       // bf16 [B*S, 3H]; this kernel writes the K and V column blocks
  // K4
  float* colsum;                 // [B, S] += sum_q mean_h P[b,h,q,k]
  float* colsum2;                // optional: queries >= colsum_split accumulate here instead
  int colsum_split;              // 0 = no split
  int colsum_valid_q;            // 1 = only valid (non-padding) queries contribute (attention_log, modeling.py:192-193)
  int pair_P, pair_chunk;        // disable_pairwise_lang_attn (model/modeling.py:160-168); pair_chunk == 0: off
  int dbg_mode;                  // timing experiments only (merlot_attention_debug_mode): 1 = no arithmetic, 2 = no MMA2, 4 = no MMA1
  unsigned long long* dbg;       // optional phase counters (merlot_attention_debug_counters): [0,8) forward, [8,16) backward
};

// phase timing of ONE softmax thread per CTA (cycles summed over all CTAs; tools/attn_phases.py prints per-tile averages)
struct PhaseClock {
  long long t, acc[8];
  bool on;
  __device__ __forceinline__ PhaseClock(bool enabled) : on(enabled) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0;
    t = on ? clock64() : 0;
  }
  __device__ __forceinline__ void lap(int phase) {
    if (on) { const long long n = clock64(); acc[phase] += n - t; t = n; }
  }
  __device__ __forceinline__ void flush(unsigned long long* dst) {
    if (on)
      for (int i = 0; i < 8; ++i) atomicAdd(dst + i, (unsigned long long)acc[i]);
  }
};

// score in the log2 domain: sc2 = scale * log2(e)
__device__ __forceinline__ float masked_score(float s, float sc2, bool vq, bool vk, bool in_range) {
  if (!in_range) return -INFINITY;
  if (!vq) return 0.0f;           // padding query row: uniform softmax (see header comment)
  return vk ? s * sc2 : -1e10f * LOG2E;
}

// -----------------------------------------------------------------------------------------------------------------
// forward (v2).  One CTA per (128-query tile, head, batch); 128 threads = one thread per TMEM lane = one query row; key tiles
// of 64 so that a CTA needs 128 TMEM columns, ~58 KB of smem and ~64 live score registers per thread: THREE CTAs are resident
// per SM and one CTA's tensor phase (S = Q K^T, O += P V) runs under the others' softmax phase.
// Per key tile:  S (tcgen05) -> the S row is read ONCE into registers -> masked online softmax with register bitmasks for key
// validity -> P (bf16) into a swizzled smem A tile -> O += P V accumulated IN TMEM.  The running maximum only advances when it
// grows by more than 2^8 (lazy rescale: O is then corrected in TMEM with tcgen05.ld/st), so the common iteration never touches
// O.  Warps whose 32 query rows all lie past the sequence end (ragged last query tile) skip the softmax arithmetic.
// -----------------------------------------------------------------------------------------------------------------
constexpr int FK = 64;  // keys per tile
constexpr int FWD_SMEM = 16384 + 8192 + 2 * 8192 + 16384 + 512 + 128 + 1024;  // Q, K, 2 x V, P, masks, barriers, alignment
constexpr int MAX_MASK_WORDS = 128;  // key-validity bitmask for up to 4096 keys
constexpr float MASKED_LOG2 = -1e10f * LOG2E;

// bitmask of in-range keys for the 32-key word starting at key k
__device__ __forceinline__ uint32_t range_word(int k, int S) {
  const int n = S - k;
  return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
}

// bits of the 32-position word starting at x0 whose positions lie in [a, b)
__device__ __forceinline__ uint32_t span_word(int x0, int a, int b) {
  const int lo = max(a - x0, 0), hi = min(b - x0, 32);
  if (hi <= lo) return 0u;
  const uint32_t below_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
  return below_hi & (0xffffffffu << lo);
}
// disable_pairwise_lang_attn (model/modeling.py:160-168): segment 0 = the P vision tokens, segment 1 + c = language chunk c;
// two positions exchange attention iff they share a segment or either is a vision token.  The relation is symmetric, so one
// helper serves "keys a query may see" (K2) and "queries a key is seen by" (K3, K4).  pair_lo_of: start of the language chunk
// of position t, or -1 when t is unrestricted (vision token / feature off); pair_word: the partners of such a position
// inside the 32-position word starting at x0.
__device__ __forceinline__ int pair_lo_of(int t, int P, int chunk) {
  return (chunk > 0 && t >= P) ? P + ((t - P) / chunk) * chunk : -1;
}
__device__ __forceinline__ uint32_t pair_word(int x0, int lo, int P, int chunk) {
  return lo < 0 ? 0xffffffffu : (span_word(x0, 0, P) | span_word(x0, lo, lo + chunk));
}

template <bool HAS_MASK>
__global__ void __launch_bounds__(128, 3) attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                                                          const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sV = smem + 24576;  // two V tiles: tile j lives in buffer j & 1 (loaded a whole tile ahead)
  uint8_t* sP = smem + 40960;  // [128 q rows][64 keys] bf16: one K-major 128B-swizzled atom
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + 57344);  // [MAX_MASK_WORDS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 57344 + 512);
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 2 /* [2] */, *bar_s = bars + 4, *bar_o = bars + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // tcgen05.mma / TMA are issued by ONE thread; the predicate must come from elect.sync (and every operand must be provably
  // warp-uniform), otherwise ptxas wraps each instruction in a per-lane ELECT / BRA.U.ANY loop that costs ~100 cycles per MMA
  const bool leader = elect_one(), warp0 = __shfl_sync(0xffffffffu, warp, 0) == 0;
  const int q0 = blockIdx.x * AT_M, h = blockIdx.y, b = blockIdx.z;
  pdl_launch_dependents();
  const int S = p.S, H = p.H;
  const int tok0 = b * S;
  const int n_kv = (S + FK - 1) / FK;

  if (tid == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv);
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(&bar_v[0], 1); mbar_init(&bar_v[1], 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_ptr, 128); tmem_relinquish(); }
  pdl_wait();
  if (HAS_MASK) {  // key validity of this batch element as a bitmask (bit k%32 of word k/32)
    for (int k = tid; k < n_kv * FK; k += 128) {
      const bool v = (k < S) ? (p.valid[tok0 + k] != 0) : false;
      const uint32_t w = __ballot_sync(0xffffffffu, v);
      if (lane == 0) s_mask[k >> 5] = w;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr, 0);  // provably warp-uniform: no per-lane waterfall around tcgen05.mma
  const uint32_t tS = tmem, tO = tmem + 64;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  if (warp0) if (leader) {
    mbar_arrive_expect_tx(bar_q, 16384);
    tma_load_2d(sQ, &tm_q, bar_q, h * AT_D, tok0 + q0);
    mbar_arrive_expect_tx(bar_k, 8192);
    tma_load_2d(sK, &tm_kv, bar_k, H + h * AT_D, tok0);
    mbar_arrive_expect_tx(&bar_v[0], 8192);
    tma_load_2d(sV, &tm_kv, &bar_v[0], 2 * H + h * AT_D, tok0);
  }

  const int q = q0 + tid;
  const bool q_in = q < S;
  const bool warp_live = (q0 + warp * 32) < S;  // at least one real query row in this warp
  const bool vq = (HAS_MASK && q_in) ? (p.valid[tok0 + q] != 0) : true;
  // a padding QUERY row softmaxes uniformly over all in-range keys: realised as zero scores with every key "valid"
  const float sc2 = vq ? p.scale * LOG2E : 0.f;
  const int pair_lo = (HAS_MASK && vq) ? pair_lo_of(q, p.pair_P, p.pair_chunk) : -1;  // a padding query row stays uniform over ALL keys
  float m_used = -INFINITY, l_run = 0.f;

  constexpr uint32_t idesc_o = make_idesc_bf16(AT_M, AT_D, 0, 1);  // B = V tile, MN-major (rows are keys)
  // a ragged last key tile (S = 266: 10 keys) only costs its 16-key units: S = Q K^T with N = 16 nu, nu k16-steps of P V
  auto nu_of = [&](int j) { return min(FK / 16, (S - j * FK + 15) >> 4); };
  auto issue_s = [&](int j) {  // S_j = Q K_j^T (tid 0); its commit also covers every MMA issued before it
    mbar_wait(bar_k, (uint32_t)(j & 1));
    tc_fence_after();
    const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK);
    const uint32_t idesc_s = make_idesc_bf16(AT_M, nu_of(j) * 16, 0, 0);
#pragma unroll
    for (int k = 0; k < AT_D / 16; ++k) umma_bf16_ss(tS, desc_kmajor(qa, k), desc_kmajor(ka, k), idesc_s, k > 0);
    umma_commit(bar_s);
  };
  if (warp0) if (leader) {
    mbar_wait(bar_q, 0);
    issue_s(0);
  }

  // Per key tile the threads wait ONCE, for { P V of tile j-1, S of tile j } (S_j is issued right behind P V_{j-1}, so the
  // tensor pipe runs them back to back); K_{j+1} and V_{j+1} are fetched while the softmax of tile j runs.
  PhaseClock pc(p.dbg != nullptr && tid == 32);
  for (int j = 0; j < n_kv; ++j) {
    const uint32_t ph = j & 1;
    const int k0 = j * FK;
    const int nu = nu_of(j);
    mbar_wait(bar_s, ph);
    tc_fence_after();
    pc.lap(0);
    if (warp0) if (leader && j + 1 < n_kv) {  // S_j and P V_{j-1} have completed: the K tile and V buffer (j+1)&1 are free
      mbar_arrive_expect_tx(bar_k, 8192);
      tma_load_2d(sK, &tm_kv, bar_k, H + h * AT_D, tok0 + (j + 1) * FK);
      mbar_arrive_expect_tx(&bar_v[(j + 1) & 1], 8192);
      tma_load_2d(sV + ((j + 1) & 1) * 8192, &tm_kv, &bar_v[(j + 1) & 1], 2 * H + h * AT_D, tok0 + (j + 1) * FK);
    }
    if (warp_live) {
      // ---- S row -> registers (log2 domain), masked; row maximum ----
      float x[FK];
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // four independent max / sum chains instead of a 64-long one
#pragma unroll
      for (int c = 0; c < FK / 32; ++c) {
        if (c * 2 >= nu) break;
        uint32_t r[32];
        tmem_ld_32x32(tS + lane_off + c * 32, r);  // (columns past N = 16 nu hold stale scores: masked to -inf below)
        tmem_wait_ld();
        const uint32_t iw = range_word(k0 + c * 32, S);
        uint32_t vw = (HAS_MASK && vq) ? s_mask[(k0 >> 5) + c] : 0xffffffffu;
        if (HAS_MASK && pair_lo >= 0) vw &= pair_word(k0 + c * 32, pair_lo, p.pair_P, p.pair_chunk);
        if (iw == 0xffffffffu && vw == 0xffffffffu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            x[c * 32 + i] = __uint_as_float(r[i]) * sc2;
            mx4[i & 3] = fmaxf(mx4[i & 3], x[c * 32 + i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float t = __uint_as_float(r[i]) * sc2;
            t = ((vw >> i) & 1u) ? t : MASKED_LOG2;
            t = ((iw >> i) & 1u) ? t : -INFINITY;
            x[c * 32 + i] = t;
            mx4[i & 3] = fmaxf(mx4[i & 3], t);
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      pc.lap(1);
      // ---- lazy running max: advance (and correct O in TMEM) only when the row max grew by more than 2^8 ----
      const bool need = mx > m_used + 8.0f;  // always true on the first tile (m_used = -inf)
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        const float f = need ? ex2_approx(m_used - mx) : 1.0f;
#pragma unroll
        for (int c = 0; c < AT_D / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(tO + lane_off + c * 32, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
          tmem_st_32x32(tO + lane_off + c * 32, r);
        }
        tmem_wait_st();
        l_run *= f;
      }
      if (need) m_used = mx;
      pc.lap(2);
      // ---- p = 2^(x - m), P (bf16) into the K-major swizzled A tile ----
      float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < FK / 32; ++c) {
        if (c * 2 >= nu) break;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = ex2_approx(x[c * 32 + i] - m_used), p1 = ex2_approx(x[c * 32 + i + 1] - m_used);
          rs4[(i >> 1) & 3] += p0 + p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(sP + sw128_offset(tid, (uint32_t)(c * 4 + g))) =
              make_uint4(pk[g * 4 + 0], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
      }
      l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
    }
    pc.lap(3);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    pc.lap(4);
    if (warp0) if (leader) {
      tc_fence_after();
      mbar_wait(&bar_v[j & 1], (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV + (j & 1) * 8192);
#pragma unroll
      for (int k = 0; k < FK / 16; ++k) {
        if (k >= nu) break;
        umma_bf16_ss(tO, desc_kmajor(pa, k), desc_mnmajor(va, k, 0), idesc_o, (j > 0 || k > 0) ? 1u : 0u);
      }
      if (j + 1 < n_kv) issue_s(j + 1);  // commits bar_s: P V_j and S_{j+1}
      else umma_commit(bar_o);
    }
  }
  mbar_wait(bar_o, 0);  // last P V: O is final
  tc_fence_after();
  pc.lap(5);
  pc.acc[6] = n_kv;

  if (warp_live) {
    const float inv = 1.0f / l_run;
    bf16* dst = p.ctx + (size_t)(tok0 + q) * p.ld_ctx + h * AT_D;
#pragma unroll
    for (int c = 0; c < AT_D / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tO + lane_off + c * 32, r);
      tmem_wait_ld();
      if (q_in) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk = make_uint4(pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * inv, __uint_as_float(r[g * 8 + 1]) * inv),
                                pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * inv, __uint_as_float(r[g * 8 + 3]) * inv),
                                pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * inv, __uint_as_float(r[g * 8 + 5]) * inv),
                                pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * inv, __uint_as_float(r[g * 8 + 7]) * inv));
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = pk;
        }
      }
    }
    if (q_in && p.lse) p.lse[((size_t)b * p.heads + h) * S + q] = (m_used + log2f(l_run)) * LN2;
  }
  pc.lap(7);
  pc.flush(p.dbg);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

// -----------------------------------------------------------------------------------------------------------------
// backward (v4): PERSISTENT, warp-specialised, software-pipelined.  One CTA per SM (320 threads) walks the work items
// (128-key tile, head, batch) assigned to it round-robin; the pipeline never drains between items:
//   warps 0-7  softmax / gradient arithmetic (two warps per TMEM lane quadrant; keys sit on the lanes),
//   warp  8    MMA issuer: every tcgen05.mma and commit (one elected lane) -- nothing else sits in its dependency chain,
//   warp  9    loader: TMA loads (K/V per item, Q/dO chunk ring) and the per-chunk statistics ring (all 32 lanes).
// The queries of an item are walked in chunks of 64; chunks are numbered globally (g) across the CTA's items.  Per chunk g:
//   MMA1(g):  S^T = K Q^T, dP^T = V dO^T                      -> TMEM buffers g & 1 (issued two chunks ahead, across items)
//   math(g):  P^T = exp2(S^T sc - lse), dS'^T = P^T (dP^T - D) -> bf16, 128B-swizzled smem buffers g & 1
//   MMA2(g):  dV += P^T dO, dK += dS'^T Q (TMEM, whole key tile), dQ_chunk = dS' K (M = 64 accumulator, TMEM buffers g & 1)
//   drain(g): dQ_chunk TMEM (16 live lanes) -> fp32 stores of whole sectors
// K/V tiles are double-buffered per item, the Q/dO chunks run through a 4-stage ring; the only CTA-wide synchronisation is at
// kernel start / end (mbarriers between the issuer and the 8 arithmetic warps, one named barrier when the batch element --
// and with it the query-validity bitmask -- changes).  At an item's end the arithmetic warps read dK / dV out of TMEM while the
// tensor pipe already runs the next item's S^T / dP^T.  (Measured on the non-persistent predecessor: 55 us of a 90 us launch
// were per-CTA prologue / epilogue / pipeline fill -- profiles/r02_attn_bwd_v3_mode_experiments.txt.)
// 1/sqrt(d) is applied once per output (dK in the epilogue, dQ in attn_dqkv_finish) instead of once per score.
// dQ never touches an atomic when the sequence has <= 4 key tiles: every item stores its partial for its key tile into its own
// slice of the [parts][tokens][H] fp32 workspace and attn_dqkv_finish sums the slices (bitwise reproducible); longer
// sequences red.add into one slice.  Warps whose 32 keys all lie past the sequence end (ragged last key tile: S = 266 has 10
// keys there) zero their P^T / dS^T rows once per item and skip the arithmetic.
// -----------------------------------------------------------------------------------------------------------------
constexpr int BQ = 64;              // queries per chunk
constexpr int MAX_DQ_PARTS = 4;     // key tiles per sequence for which dQ goes through per-tile slices instead of atomics
constexpr int BWD_QSTAGES = 5;      // Q/dO chunk ring: the loader runs up to 5 chunks ahead of MMA2, 3 ahead of MMA1 (load latency ~1 us)
constexpr int BWD_THREADS = 320;
constexpr int BWD_STSTAGES = 4;     // statistics ring (chunks)
constexpr int BWD_SMEM = 2 * 32768 + BWD_QSTAGES * 16384 + 2 * 16384 + 2 * 16384 + 2048 + 512 + 512 + 1024;

template <bool HAS_MASK, bool DQ_ATOMIC>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_q,
                const __grid_constant__ CUtensorMap tm_do, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sKV = smem;                                       // [2] x { K tile 16 KB, V tile 16 KB }
  uint8_t* sQd = smem + 2 * 32768;                           // BWD_QSTAGES x { Q chunk 8 KB, dO chunk 8 KB }
  uint8_t* sPT = sQd + BWD_QSTAGES * 16384;                  // [2] P^T  [128 keys][64 q] bf16: one 128B-swizzled atom each
  uint8_t* sdST = sPT + 2 * 16384;                           // [2] dS^T same layout
  // (the 8 warp-private 2 KB transposition slots of the dK/dV epilogue live in the P^T buffer that the NEXT item's first chunk
  //  does not use: every MMA of the item has completed by then, and no warp can reach the chunk after that one -- it needs all
  //  eight warps' arrivals for the first -- before every warp has left its epilogue)
  float* s_nlse = reinterpret_cast<float*>(sdST + 2 * 16384);  // [BWD_STSTAGES][64]  -lse * log2(e)
  float* s_dsum = s_nlse + BWD_STSTAGES * BQ;                  // [BWD_STSTAGES][64]
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_dsum + BWD_STSTAGES * BQ);  // [MAX_MASK_WORDS] query validity bits of the current batch element
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_mask) + 512);
  uint64_t *bar_kv = bars /* [2] full */, *bar_kvfree = bars + 2 /* [2] */, *bar_q = bars + 4 /* [<= 6] full */, *bar_qfree = bars + 10 /* [<= 6] */,
           *bar_st = bars + 16 /* [4] full */, *bar_stfree = bars + 20 /* [4] */, *bar_s = bars + 24 /* [2] */, *bar_p = bars + 26 /* [2] */,
           *bar_d = bars + 28 /* [2] */, *bar_acc = bars + 30;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 31);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform (see attn_fwd_kernel)
  const bool leader = elect_one();
  const int wg = (warp >> 2) & 1, quad = warp & 3, row_t = quad * 32 + lane;
  pdl_launch_dependents();
  const int S = p.S, H = p.H;
  const int n_q = (S + BQ - 1) / BQ;           // chunks per item
  const int n_kv = (S + AT_N - 1) / AT_N;      // key tiles per (head, batch)
  const int total_items = n_kv * p.heads * p.B;
  const int n_local = ((int)blockIdx.x < total_items) ? (total_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int G = n_local * n_q;                 // chunks this CTA walks

  if (tid == 0) {
    tma_prefetch_desc(&tm_kv); tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_kv[i], 1); mbar_init(&bar_kvfree[i], 1); }
    for (int i = 0; i < BWD_QSTAGES; ++i) { mbar_init(&bar_q[i], 1); mbar_init(&bar_qfree[i], 1); }
    for (int i = 0; i < BWD_STSTAGES; ++i) { mbar_init(&bar_st[i], 1); mbar_init(&bar_stfree[i], 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 8); mbar_init(&bar_d[i], 1); }
    mbar_init(bar_acc, 8);
    fence_barrier_init();
  }
  if (warp == 8) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  // TMEM columns: S^T[2] 0/64, dP^T[2] 128/192, dV 256, dK 320, dQ[2] 384/448
  const uint32_t tdV = tmem + 256, tdK = tmem + 320;
  // a ragged last chunk only costs its 16-query units (S^T / dP^T with N = 16 nu, nu arithmetic units, nu k16-steps of dV / dK)
  auto nu_of = [&](int c) { return min(BQ / 16, (S - c * BQ + 15) >> 4); };
  // item `it` of this CTA -> (key tile, head, batch)
  auto decode = [&](int it, int& kt, int& hh, int& bb_) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    kt = item % n_kv;
    const int rest = item / n_kv;
    hh = rest % p.heads;
    bb_ = rest / p.heads;
  };

  // Chunk cursors: (item, chunk-in-item), ring stages and their parities advanced incrementally -- an integer division per chunk
  // in a single warp's issue path costs more than the MMAs it feeds (measured); the item decode runs once per item.
  struct Cur {
    int g, it, c, kt, h, b;
    int st;        // Q/dO ring stage of chunk g
    uint32_t sph;  // parity of that stage's current use
  };
  auto cur_init = [&](Cur& cu) { cu.g = 0; cu.it = 0; cu.c = 0; cu.st = 0; cu.sph = 0; decode(0, cu.kt, cu.h, cu.b); };
  auto cur_next = [&](Cur& cu) {
    ++cu.g;
    if (++cu.st == BWD_QSTAGES) { cu.st = 0; cu.sph ^= 1u; }
    if (++cu.c == n_q) { cu.c = 0; ++cu.it; if (cu.it < n_local) decode(cu.it, cu.kt, cu.h, cu.b); }
  };

  if (warp == 9) {
    // ================================= loader warp =================================
    // K/V tiles per item (2 buffers), Q/dO chunks (ring), statistics (-lse*log2e and D, ring of 4 chunks; the global loads are
    // issued two chunks before the values are parked in smem).  Classic producer: waits on the "free" barrier of a slot with
    // the inverted parity, so the first pass through a ring never blocks.
    Cur cl, cf;
    cur_init(cl); cur_init(cf);
    float st_nl[2][2], st_ds[2][2];
    auto fetch_stats = [&](const Cur& cu) {
      const size_t o0 = ((size_t)cu.b * p.heads + cu.h) * S;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int q = cu.c * BQ + lane + r * 32;
        const float nl = (q < S) ? p.lse[o0 + q] : 0.f, dd = (q < S) ? p.dsum[o0 + q] : 0.f;
        if (cu.g & 1) { st_nl[1][r] = nl; st_ds[1][r] = dd; } else { st_nl[0][r] = nl; st_ds[0][r] = dd; }
      }
    };
    for (int i = 0; i < 2 && i < G; ++i) { fetch_stats(cf); cur_next(cf); }
    for (int L = 0; L < G; ++L) {
      if (cl.c == 0) {  // first chunk of item cl.it: its K/V tiles
        const int it = cl.it;
        mbar_wait(&bar_kvfree[it & 1], (uint32_t)(((it >> 1) & 1) ^ 1));
        if (leader) {
          uint8_t* buf = sKV + (it & 1) * 32768;
          mbar_arrive_expect_tx(&bar_kv[it & 1], 32768);
          tma_load_2d(buf, &tm_kv, &bar_kv[it & 1], H + cl.h * AT_D, cl.b * S + cl.kt * AT_N);
          tma_load_2d(buf + 16384, &tm_kv, &bar_kv[it & 1], 2 * H + cl.h * AT_D, cl.b * S + cl.kt * AT_N);
        }
      }
      mbar_wait(&bar_qfree[cl.st], cl.sph ^ 1u);
      if (leader) {
        uint8_t* buf = sQd + cl.st * 16384;
        mbar_arrive_expect_tx(&bar_q[cl.st], 16384);
        tma_load_2d(buf, &tm_q, &bar_q[cl.st], cl.h * AT_D, cl.b * S + cl.c * BQ);
        tma_load_2d(buf + 8192, &tm_do, &bar_q[cl.st], cl.h * AT_D, cl.b * S + cl.c * BQ);
      }
      {  // statistics of chunk L (fetched two iterations ago) -> ring slot L & 3
        const int sb = L & (BWD_STSTAGES - 1);
        mbar_wait(&bar_stfree[sb], (uint32_t)(((L / BWD_STSTAGES) & 1) ^ 1));
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          s_nlse[sb * BQ + lane + r * 32] = -((L & 1) ? st_nl[1][r] : st_nl[0][r]) * LOG2E;
          s_dsum[sb * BQ + lane + r * 32] = (L & 1) ? st_ds[1][r] : st_ds[0][r];
        }
        __syncwarp();
        if (leader) mbar_arrive(&bar_st[sb]);
        if (L + 2 < G) { fetch_stats(cf); cur_next(cf); }
      }
      cur_next(cl);
    }
  } else if (warp == 8) {
    // ================================= MMA issuer warp =================================
    constexpr uint32_t idesc_dv = make_idesc_bf16(AT_N, AT_D, 0, 1);   // A = P^T/dS^T K-major (q), B = dO/Q MN-major
    constexpr uint32_t idesc_dq = make_idesc_bf16(BQ, AT_D, 1, 1);     // M = 64: A = dS (MN-major view of dS^T), B = K MN-major
    int kv_ready = -1;  // last item whose K/V tiles this warp has seen land
    auto issue_st = [&](const Cur& cu) {  // MMA1: S^T and dP^T of a chunk -> TMEM buffers g & 1
      const int g = cu.g, it = cu.it;
      if (it > kv_ready) { mbar_wait(&bar_kv[it & 1], (uint32_t)((it >> 1) & 1)); kv_ready = it; }
      mbar_wait(&bar_q[cu.st], cu.sph);
      tc_fence_after();
      if (leader) {
        const uint32_t ka = smem_u32(sKV + (it & 1) * 32768), va = ka + 16384;
        const uint32_t qa = smem_u32(sQd + cu.st * 16384), da = qa + 8192;
        const uint32_t idesc_st = make_idesc_bf16(AT_N, nu_of(cu.c) * 16, 0, 0);  // both operands K-major (d)
        const uint32_t tST = tmem + (g & 1) * 64, tdPT = tmem + 128 + (g & 1) * 64;
        if (!(p.dbg_mode & 4)) {
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) umma_bf16_ss(tST, desc_kmajor(ka, k), desc_kmajor(qa, k), idesc_st, k > 0);
#pragma unroll
          for (int k = 0; k < AT_D / 16; ++k) umma_bf16_ss(tdPT, desc_kmajor(va, k), desc_kmajor(da, k), idesc_st, k > 0);
        }
        umma_commit(&bar_s[g & 1]);
      }
    };
    Cur c_m1, c_m2;  // next chunk to run MMA1 / MMA2 on
    cur_init(c_m1); cur_init(c_m2);
    if (G > 0) { issue_st(c_m1); cur_next(c_m1); }
    if (G > 1) { issue_st(c_m1); cur_next(c_m1); }
    PhaseClock ic(p.dbg != nullptr && leader);
    for (int g = 0; g < G; ++g) {
      const int bb = g & 1;
      const int it = c_m2.it, c = c_m2.c;
      const int nu = nu_of(c);
      if (c == 0 && it > 0) mbar_wait(bar_acc, (uint32_t)((it - 1) & 1));  // dK / dV of the previous item have left TMEM
      mbar_wait(&bar_p[bb], (uint32_t)((g >> 1) & 1));  // P^T / dS^T of chunk g are in smem; its S^T / dP^T buffers are free;
      tc_fence_after();                                 // dQ[bb] of chunk g-2 has been drained (program order per warp)
      ic.lap(0);
      if (leader) {
        const uint32_t pa = smem_u32(sPT + bb * 16384), sa = smem_u32(sdST + bb * 16384);
        const uint32_t qa = smem_u32(sQd + c_m2.st * 16384), da = qa + 8192;
        const uint32_t ka = smem_u32(sKV + (it & 1) * 32768);
        const uint32_t tdQ = tmem + 384 + bb * 64;
        if (!(p.dbg_mode & 2)) {
#pragma unroll
          for (int k = 0; k < BQ / 16; ++k)  // dV += P^T dO   (contraction over q)
            if (k < nu) umma_bf16_ss(tdV, desc_kmajor(pa, k), desc_mnmajor(da, k, 0), idesc_dv, (c > 0 || k > 0));
#pragma unroll
          for (int k = 0; k < BQ / 16; ++k)  // dK += dS^T Q
            if (k < nu) umma_bf16_ss(tdK, desc_kmajor(sa, k), desc_mnmajor(qa, k, 0), idesc_dv, (c > 0 || k > 0));
          const int nk = min(AT_N / 16, (S - c_m2.kt * AT_N + 15) >> 4);  // a ragged last key tile contracts over its keys only
#pragma unroll
          for (int k = 0; k < AT_N / 16; ++k)  // dQ_chunk = dS K  (contraction over the keys; stale columns of a ragged chunk
            if (k < nk) umma_bf16_ss(tdQ, desc_mnmajor(sa, k, 0), desc_mnmajor(ka, k, 0), idesc_dq, k > 0);  // only feed rows never stored)
        }
        umma_commit(&bar_d[bb]);               // dQ[bb] ready; P^T / dS^T[bb] free
        umma_commit(&bar_qfree[c_m2.st]);      // this chunk's Q/dO ring stage is free
        if (c == n_q - 1) umma_commit(&bar_kvfree[it & 1]);  // the item's K/V buffer is free
      }
      ic.lap(1);
      if (g + 2 < G) { issue_st(c_m1); cur_next(c_m1); }
      ic.lap(3);
      cur_next(c_m2);
    }
    ic.acc[6] = G;
    ic.flush(p.dbg ? p.dbg + 16 : nullptr);
  } else {
    // ================================= arithmetic warps =================================
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float sc2 = p.scale * LOG2E;
    PhaseClock pc(p.dbg != nullptr && tid == 32);
    int cur_b = -1;
    for (int it = 0; it < n_local; ++it) {
      int kt, h, b;
      decode(it, kt, h, b);
      const int k0 = kt * AT_N, tok0 = b * S;
      if (HAS_MASK && b != cur_b) {  // query-validity bitmask of this batch element (every warp is past the previous item here)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int k = tid; k < n_q * BQ; k += 256) {
          const bool v = (k < S) ? (p.valid[tok0 + k] != 0) : false;
          const uint32_t w = __ballot_sync(0xffffffffu, v);
          if (lane == 0) s_mask[k >> 5] = w;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      cur_b = b;
      const int kk = k0 + row_t;  // this thread's key row
      const bool k_in = kk < S;
      const bool vk = k_in ? (HAS_MASK ? p.valid[tok0 + kk] != 0 : true) : false;
      const int pair_lo = HAS_MASK ? pair_lo_of(kk, p.pair_P, p.pair_chunk) : -1;  // queries this key is seen by
      // a warp whose 32 keys are all out of range contributes exact zeros: written once per item (both buffers: every MMA of
      // the previous item has completed -- its last drain waited for that), arithmetic skipped afterwards
      const bool warp_dead = (k0 + quad * 32) >= S || (p.dbg_mode & 1);
      if (warp_dead && wg == 0) {
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
#pragma unroll
          for (int b2 = 0; b2 < 2; ++b2) {
            *reinterpret_cast<uint4*>(sPT + b2 * 16384 + sw128_offset(row_t, c8)) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(sdST + b2 * 16384 + sw128_offset(row_t, c8)) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
      }
      float* dq_base = p.dq_accum + (DQ_ATOMIC ? (size_t)0 : (size_t)kt * p.dq_part_stride);

      // dQ partial of chunk c (global index g): M = 64 accumulator, row 16*quad + l on lane 32*quad + l (l < 16); this warp owns
      // columns [32 wg, 32 wg + 32).  16x256b: only the 16 live lanes are read; thread t holds rows t/4 and t/4 + 8, two adjacent
      // columns per 8-column block: every store instruction writes 8 rows x 32 B (whole sectors).
      auto drain = [&](int c, int g) {
        mbar_wait(&bar_d[g & 1], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        pc.lap(3);
        uint32_t r[16];
        tmem_ld_16x256b_x4(tmem + 384 + (g & 1) * 64 + lane_off + wg * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int q = c * BQ + quad * 16 + (lane >> 2) + hh * 8;
          if (q < S) {
            float* dst = dq_base + (size_t)(tok0 + q) * p.ld_dq + h * AT_D + wg * 32 + (lane & 3) * 2;
#pragma unroll
            for (int jb = 0; jb < 4; ++jb) {
              const float v0 = __uint_as_float(r[jb * 4 + hh * 2]), v1 = __uint_as_float(r[jb * 4 + hh * 2 + 1]);
              if (DQ_ATOMIC) {
                asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + jb * 8), "f"(v0), "f"(v1) : "memory");
              } else {
                *reinterpret_cast<float2*>(dst + jb * 8) = make_float2(v0, v1);
              }
            }
          }
        }
        pc.lap(4);
      };

      for (int c = 0; c < n_q; ++c) {
        const int g = it * n_q + c;
        const int bb = g & 1;
        const int q0 = c * BQ;
        const int nu = nu_of(c);
        const int sb = g & (BWD_STSTAGES - 1);
        mbar_wait(&bar_st[sb], (uint32_t)((g / BWD_STSTAGES) & 1));
        mbar_wait(&bar_s[bb], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        pc.lap(0);
        if (!warp_dead) {
          const float* nlse = s_nlse + sb * BQ;
          const float* dsm = s_dsum + sb * BQ;
          uint8_t* pT = sPT + bb * 16384;
          uint8_t* dT = sdST + bb * 16384;
          // the TMEM loads of the second unit run under the arithmetic of the first (TMEM -> register bandwidth is ~64 B/clk per
          // SM: S^T and dP^T of one chunk are 64 KB)
          uint32_t rs[16], rd[16], rs2[16], rd2[16];
          const int u0 = wg * 2;
          if (u0 < nu) {
            tmem_ld_32x16(tmem + bb * 64 + lane_off + u0 * 16, rs);
            tmem_ld_32x16(tmem + 128 + bb * 64 + lane_off + u0 * 16, rd);
            tmem_wait_ld_regs16x2(rs, rd);
          }
          if (u0 + 1 < nu) {
            tmem_ld_32x16(tmem + bb * 64 + lane_off + (u0 + 1) * 16, rs2);
            tmem_ld_32x16(tmem + 128 + bb * 64 + lane_off + (u0 + 1) * 16, rd2);
          }
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            const int u = u0 + uu;
            if (u >= nu) break;  // warp-uniform
            if (uu == 1) {
              tmem_wait_ld_regs16x2(rs2, rd2);  // the registers are threaded through the wait: no use can be scheduled above it
#pragma unroll
              for (int e = 0; e < 16; ++e) { rs[e] = rs2[e]; rd[e] = rd2[e]; }
            }
            const int qb = q0 + u * 16;
            const uint32_t iw = k_in ? ((range_word(qb & ~31, S) >> (qb & 31)) & 0xffffu) : 0u;   // query in range (and this key in range)
            const uint32_t qw = HAS_MASK ? ((s_mask[qb >> 5] >> (qb & 31)) & 0xffffu) : 0xffffu;   // query validity
            const uint32_t aw = (HAS_MASK && pair_lo >= 0) ? ((pair_word(qb & ~31, pair_lo, p.pair_P, p.pair_chunk) >> (qb & 31)) & 0xffffu) : 0xffffu;
            const bool fast = (iw == 0xffffu) && (qw == 0xffffu) && vk && (aw == 0xffffu);
            uint32_t pk[8], dk[8];
#pragma unroll
            for (int e = 0; e < 16; e += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(nlse + u * 16 + e);
              const float4 d4 = *reinterpret_cast<const float4*>(dsm + u * 16 + e);
              const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, ds[4] = {d4.x, d4.y, d4.z, d4.w};
              float pv[4], dv[4];
#pragma unroll
              for (int t4 = 0; t4 < 4; ++t4) {
                float pr;
                if (fast) {
                  pr = ex2_approx(fmaf(__uint_as_float(rs[e + t4]), sc2, ls[t4]));
                } else {
                  float t = __uint_as_float(rs[e + t4]) * sc2;
                  t = (vk && ((aw >> (e + t4)) & 1u)) ? t : MASKED_LOG2;
                  t = ((qw >> (e + t4)) & 1u) ? t : 0.f;  // padding query: uniform row (zero scores)
                  pr = ex2_approx(t + ls[t4]);
                  pr = ((iw >> (e + t4)) & 1u) ? pr : 0.f;
                }
                pv[t4] = pr;
                // d(score)/d(q k^T) = m * scale (utils/transformer.py:109-110: scores*m - 1e10*(1-m)): a padding QUERY row keeps
                // its uniform probabilities for dV but sends nothing back into q and k.  (scale itself: dK epilogue / dQ finish)
                float gq = pr;
                if (!fast) gq = ((qw >> (e + t4)) & 1u) ? gq : 0.f;
                dv[t4] = (__uint_as_float(rd[e + t4]) - ds[t4]) * gq;
              }
              pk[(e >> 1)] = pack_bf16x2(pv[0], pv[1]); pk[(e >> 1) + 1] = pack_bf16x2(pv[2], pv[3]);
              dk[(e >> 1)] = pack_bf16x2(dv[0], dv[1]); dk[(e >> 1) + 1] = pack_bf16x2(dv[2], dv[3]);
            }
#pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
              const uint32_t off = sw128_offset(row_t, (uint32_t)(u * 2 + g2));
              *reinterpret_cast<uint4*>(pT + off) = make_uint4(pk[g2 * 4], pk[g2 * 4 + 1], pk[g2 * 4 + 2], pk[g2 * 4 + 3]);
              *reinterpret_cast<uint4*>(dT + off) = make_uint4(dk[g2 * 4], dk[g2 * 4 + 1], dk[g2 * 4 + 2], dk[g2 * 4 + 3]);
            }
          }
        }
        pc.lap(1);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bar_p[bb]); mbar_arrive(&bar_stfree[sb]); }
        pc.lap(2);
        if (c >= 1) drain(c - 1, g - 1);
      }
      drain(n_q - 1, it * n_q + n_q - 1);  // also: every MMA of this item has completed (its commit covers all earlier ones)
      // ---- dK (x 1/sqrt(d)), dV for this key tile (exclusive rows), transposed through the slot: 8 rows x 64 B per instruction ----
      uint8_t* slot = sPT + ((((it + 1) * n_q) & 1) ^ 1) * 16384 + warp * 2048;
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        uint32_t r[32];
        tmem_ld_32x32((which == 0 ? tdK : tdV) + lane_off + wg * 32, r);
        tmem_wait_ld();
        const float mul = which == 0 ? p.scale : 1.0f;
#pragma unroll
        for (int g4 = 0; g4 < 4; ++g4)  // the slot as [32 rows][64 B]: row `lane` = this thread's 32 bf16 values
          *reinterpret_cast<uint4*>(slot + lane * 64 + ((g4 ^ ((lane >> 1) & 3)) << 4)) =
              make_uint4(pack_bf16x2(__uint_as_float(r[g4 * 8 + 0]) * mul, __uint_as_float(r[g4 * 8 + 1]) * mul),
                         pack_bf16x2(__uint_as_float(r[g4 * 8 + 2]) * mul, __uint_as_float(r[g4 * 8 + 3]) * mul),
                         pack_bf16x2(__uint_as_float(r[g4 * 8 + 4]) * mul, __uint_as_float(r[g4 * 8 + 5]) * mul),
                         pack_bf16x2(__uint_as_float(r[g4 * 8 + 6]) * mul, __uint_as_float(r[g4 * 8 + 7]) * mul));
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int rr = i4 * 8 + (lane >> 2), cc = lane & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(slot + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
          const int key = k0 + quad * 32 + rr;
          if (key < S)
            *reinterpret_cast<uint4*>(p.dqkv + (size_t)(tok0 + key) * p.ld_dqkv + (which == 0 ? H : 2 * H) + h * AT_D + wg * 32 + cc * 8) = v;
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc);  // dK / dV are out of TMEM: the next item's MMA2 may overwrite them
      pc.lap(7);
    }
    pc.acc[6] = G;
    pc.flush(p.dbg ? p.dbg + 8 : nullptr);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// D[b,h,q] = sum_d dO[q,hd] * O[q,hd]   (one warp per (token, head) pair would waste lanes; 8 lanes x 8 elems per head)
__global__ void attn_dsum_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o, int ld, float* __restrict__ dsum,
                                 int B, int S, int heads) {
  pdl_launch_dependents();
  pdl_wait();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long item = gid >> 3;  // (token, head)
  const int sub = (int)(gid & 7);
  const long long total = (long long)B * S * heads;
  float acc = 0.f;
  if (item < total) {
    const int hh = (int)(item % heads);
    const long long tok = item / heads;
    const size_t off = (size_t)tok * ld + hh * AT_D + sub * 8;
    uint4 a = __ldg(reinterpret_cast<const uint4*>(o + off));
    uint4 g = __ldg(reinterpret_cast<const uint4*>(d_o + off));
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* pg = reinterpret_cast<const uint32_t*>(&g);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 x = unpack_bf16x2(pa[i]), y = unpack_bf16x2(pg[i]);
      acc += x.x * y.x + x.y * y.y;
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (item < total && sub == 0) {
    const int hh = (int)(item % heads);
    const long long tok = item / heads;
    const int bb = (int)(tok / S), q = (int)(tok % S);
    dsum[((size_t)bb * heads + hh) * S + q] = acc;
  }
}

// dq fp32 accumulator -> bf16 q-block of dqkv (and re-zero the accumulator for the next layer), fused with the column sums
// of the whole dqkv row block = the gradient of the fused q/k/v bias.  grid = (ceil(3H/256), row slabs), 8 warps, lane = 8 cols.
// n_parts > 0: dq holds n_parts per-key-tile slices (part_stride elements apart) that are summed here; n_parts == 0: one
// atomically accumulated slice that is re-zeroed for the next layer.
__global__ void __launch_bounds__(256) attn_dqkv_finish_kernel(float* __restrict__ dq, int ld_dq, bf16* __restrict__ dqkv, int ld_dqkv,
                                                               long long rows, int H, float* __restrict__ bias_grad, int n_parts,
                                                               size_t part_stride, float scale) {
  __shared__ float sred[8][256];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.x * 256 + lane * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < 3 * H) {
    for (long long r = (long long)blockIdx.y * 8 + warp; r < rows; r += (long long)gridDim.y * 8) {
      float v[8];
      bf16* o = dqkv + (size_t)r * ld_dqkv + col;
      if (col < H) {
        float4* src = reinterpret_cast<float4*>(dq + (size_t)r * ld_dq + col);
        float4 a = src[0], b = src[1];
        if (n_parts == 0) {
          src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
          src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          for (int t = 1; t < n_parts; ++t) {  // fixed order: the sum is bitwise reproducible
            const float4* s2 = reinterpret_cast<const float4*>(dq + (size_t)t * part_stride + (size_t)r * ld_dq + col);
            const float4 c = s2[0], d = s2[1];
            a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
            b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
          }
        }
        // K3 accumulates dS' K without the 1/sqrt(d) of the scores: applied here, once per output element
        const uint4 pk = make_uint4(pack_bf16x2(a.x * scale, a.y * scale), pack_bf16x2(a.z * scale, a.w * scale),
                                    pack_bf16x2(b.x * scale, b.y * scale), pack_bf16x2(b.z * scale, b.w * scale));
        *reinterpret_cast<uint4*>(o) = pk;
        const uint32_t* pu = reinterpret_cast<const uint32_t*>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf16x2(pu[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
      } else {
        const uint4 u = *reinterpret_cast<const uint4*>(o);
        const uint32_t* pu = reinterpret_cast<const uint32_t*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = unpack_bf16x2(pu[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sred[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  const int c = threadIdx.x;
  if (bias_grad != nullptr && blockIdx.x * 256 + c < 3 * H) {
    float s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s2 += sred[w][c];
    atomicAdd(bias_grad + blockIdx.x * 256 + c, s2);
  }
}

// -----------------------------------------------------------------------------------------------------------------
// K4: colsum[b,k] += (1/heads) * sum_q P[b,h,q,k], recomputed from (q,k,lse); keys on TMEM lanes so the reduction over
// queries runs along registers.  One CTA per (key tile, head, batch), two resident per SM; the query tiles are double-buffered
// in smem AND in TMEM: S^T of tile i+1 is issued before the exponentials of tile i start.
// -----------------------------------------------------------------------------------------------------------------
constexpr int CS_SMEM = 16384 * 3 + 1024 + 512 + 128 + 1024;

template <bool HAS_MASK>
__global__ void __launch_bounds__(128, 2) attn_colsum_kernel(const __grid_constant__ CUtensorMap tm_qkv, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sQ = smem + 16384;                                         // [2] x 16 KB
  float* s_nlse = reinterpret_cast<float*>(smem + 49152);             // [2][128]  -lse * log2(e)
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(smem + 49152 + 1024);  // query validity bits
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152 + 1024 + 512);
  uint64_t *bar_k = bars, *bar_q = bars + 1 /* [2] */, *bar_s = bars + 3 /* [2] */;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool leader = elect_one(), warp0 = __shfl_sync(0xffffffffu, warp, 0) == 0;  // see attn_fwd_kernel
  const int k0 = blockIdx.x * AT_N, h = blockIdx.y, b = blockIdx.z;
  const int S = p.S, H = p.H, tok0 = b * S;
  const int n_q = (S + AT_M - 1) / AT_M;
  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(bar_k, 1); mbar_init(&bar_q[0], 1); mbar_init(&bar_q[1], 1); mbar_init(&bar_s[0], 1); mbar_init(&bar_s[1], 1);
    fence_barrier_init();
  }
  pdl_launch_dependents();
  if (warp == 0) { tmem_alloc(tmem_ptr, 256); tmem_relinquish(); }
  pdl_wait();
  if (HAS_MASK) {
    for (int k = tid; k < n_q * AT_M; k += 128) {
      const bool v = (k < S) ? (p.valid[tok0 + k] != 0) : false;
      const uint32_t w = __ballot_sync(0xffffffffu, v);
      if (lane == 0) s_mask[k >> 5] = w;
    }
  }
  auto stage_lse = [&](int i) {
    const int q = i * AT_M + tid;
    s_nlse[(i & 1) * AT_M + tid] = (q < S) ? -p.lse[((size_t)b * p.heads + h) * S + q] * LOG2E : 0.f;
  };
  stage_lse(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr, 0);  // provably warp-uniform: no per-lane waterfall around tcgen05.mma
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  auto load_q = [&](int i) {
    mbar_arrive_expect_tx(&bar_q[i & 1], 16384);
    tma_load_2d(sQ + (i & 1) * 16384, &tm_qkv, &bar_q[i & 1], h * AT_D, tok0 + i * AT_M);
  };
  auto ncq_of = [&](int i) { return min(AT_M / 32, (S - i * AT_M + 31) >> 5); };
  auto issue = [&](int i) {  // S^T of query tile i -> TMEM buffer i & 1
    mbar_wait(&bar_q[i & 1], (uint32_t)((i >> 1) & 1));
    tc_fence_after();
    const uint32_t ka = smem_u32(sK), qa = smem_u32(sQ + (i & 1) * 16384);
    const uint32_t idesc = make_idesc_bf16(AT_N, ncq_of(i) * 32, 0, 0);
#pragma unroll
    for (int k = 0; k < AT_D / 16; ++k) umma_bf16_ss(tmem + (i & 1) * 128, desc_kmajor(ka, k), desc_kmajor(qa, k), idesc, k > 0);
    umma_commit(&bar_s[i & 1]);
  };
  if (warp0) if (leader) {
    mbar_arrive_expect_tx(bar_k, 16384);
    tma_load_2d(sK, &tm_qkv, bar_k, H + h * AT_D, tok0 + k0);
    load_q(0);
    if (n_q > 1) load_q(1);
    mbar_wait(bar_k, 0);
    issue(0);
  }
  const int kk = k0 + tid;
  const bool k_in = kk < S;
  const bool vk = k_in ? (HAS_MASK ? p.valid[tok0 + kk] != 0 : true) : false;
  const int pair_lo = HAS_MASK ? pair_lo_of(kk, p.pair_P, p.pair_chunk) : -1;  // queries this key is seen by
  const bool warp_live = (k0 + warp * 32) < S;
  const float sc2 = p.scale * LOG2E;
  float acc = 0.f, acc2 = 0.f;
  const int split = p.colsum2 ? p.colsum_split : 0x7fffffff;
  for (int i = 0; i < n_q; ++i) {
    const int q0 = i * AT_M;
    if (warp0) if (leader && i + 1 < n_q) issue(i + 1);   // runs under the exponentials below
    if (i + 1 < n_q) stage_lse(i + 1);
    mbar_wait(&bar_s[i & 1], (uint32_t)((i >> 1) & 1));
    tc_fence_after();
    if (warp0) if (leader && i + 2 < n_q) load_q(i + 2);  // S^T_i has consumed Q buffer i & 1
    const float* nl = s_nlse + (i & 1) * AT_M;
    const int ncq = ncq_of(i);
    if (warp_live) {
#pragma unroll 1
      for (int c = 0; c < AT_M / 32; ++c) {
        if (c >= ncq) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem + (i & 1) * 128 + lane_off + c * 32, r);
        tmem_wait_ld();
        const uint32_t iw = k_in ? range_word(q0 + c * 32, S) : 0u;              // query (and this key) in range
        const uint32_t qw = HAS_MASK ? s_mask[(q0 >> 5) + c] : 0xffffffffu;      // query validity
        const uint32_t sw = range_word(q0 + c * 32, split);                     // query < split -> colsum, else colsum2
        const uint32_t keep = iw & ((HAS_MASK && p.colsum_valid_q) ? qw : 0xffffffffu);  // queries that contribute at all
        const uint32_t aw = (HAS_MASK && pair_lo >= 0) ? pair_word(q0 + c * 32, pair_lo, p.pair_P, p.pair_chunk) : 0xffffffffu;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        if (keep == 0xffffffffu && qw == 0xffffffffu && vk && aw == 0xffffffffu && (sw == 0xffffffffu || sw == 0u)) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            a0 += ex2_approx(fmaf(__uint_as_float(r[e]), sc2, nl[c * 32 + e]));
            a1 += ex2_approx(fmaf(__uint_as_float(r[e + 1]), sc2, nl[c * 32 + e + 1]));
          }
          if (sw) acc += a0 + a1; else acc2 += a0 + a1;
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            float x = __uint_as_float(r[e]) * sc2;
            x = (vk && ((aw >> e) & 1u)) ? x : MASKED_LOG2;
            x = ((qw >> e) & 1u) ? x : 0.f;  // padding query: uniform row (zero scores)
            float pr = ex2_approx(x + nl[c * 32 + e]);
            pr = ((keep >> e) & 1u) ? pr : 0.f;
            if ((sw >> e) & 1u) a0 += pr; else b0 += pr;
          }
          acc += a0; acc2 += b0;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM buffer i & 1 and s_nlse[i & 1] are reused two tiles later
  }
  if (k_in) {
    atomicAdd(p.colsum + (size_t)b * S + kk, acc / (float)p.heads);
    if (p.colsum2) atomicAdd(p.colsum2 + (size_t)b * S + kk, acc2 / (float)p.heads);
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

static unsigned long long* g_attn_dbg = nullptr;  // merlot_attention_debug_counters
static int g_attn_dbg_mode = 0;                   // merlot_attention_debug_mode

static int check_common(const merlot_attn_t* a) {
  MB_REQUIRE(a != nullptr, MERLOT_EINVAL, "attention: null descriptor");
  MB_REQUIRE(a->B > 0 && a->S > 0 && a->heads > 0, MERLOT_ESHAPE, "attention: bad dims B=%d S=%d heads=%d", a->B, a->S,
             a->heads);
  MB_REQUIRE(a->head_dim == 64, MERLOT_ESHAPE, "attention: head size must be 64 (got %d)", a->head_dim);
  MB_REQUIRE(a->qkv != nullptr && a->ld_qkv >= 3 * a->heads * 64 && (a->ld_qkv % 8) == 0, MERLOT_ESHAPE,
             "attention: qkv must be [tokens, >=3H] with ld %% 8 == 0");
  MB_REQUIRE(a->pair_chunk_len >= 0 && a->pair_viz_len >= 0 && (a->pair_chunk_len == 0 || a->valid != nullptr), MERLOT_EINVAL,
             "attention: pair_chunk_len > 0 (disable_pairwise_lang_attn) needs the token-validity mask and non-negative lengths");
  return MERLOT_OK;
}

static void fill_dev(const merlot_attn_t* a, AttnDev* p) {
  memset(p, 0, sizeof(*p));
  p->B = a->B; p->S = a->S; p->heads = a->heads; p->H = a->heads * 64;
  p->valid = reinterpret_cast<const uint8_t*>(a->valid);
  p->scale = a->scale;
  p->ctx = reinterpret_cast<bf16*>(a->ctx); p->ld_ctx = a->ld_ctx;
  p->lse = a->lse;
  p->dsum = a->dsum;
  p->dq_accum = a->dq_accum; p->ld_dq = a->ld_dq;
  p->dqkv = reinterpret_cast<bf16*>(a->dqkv); p->ld_dqkv = a->ld_dqkv;
  p->colsum = a->colsum;
  p->colsum2 = a->colsum2;
  p->colsum_split = a->colsum_split;
  p->colsum_valid_q = a->colsum_valid_q;
  p->pair_P = a->pair_viz_len; p->pair_chunk = a->pair_chunk_len;
  p->dbg = g_attn_dbg;
  p->dbg_mode = g_attn_dbg_mode;
}

}  // namespace mb

using namespace mb;

extern "C" void merlot_attention_debug_counters(void* buf_u64x16) { g_attn_dbg = reinterpret_cast<unsigned long long*>(buf_u64x16); }

extern "C" void merlot_attention_debug_mode(int mode) { g_attn_dbg_mode = mode; }

extern "C" int merlot_attention_fwd(const merlot_attn_t* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_common(a);
  if (rc) return rc;
  MB_REQUIRE(a->ctx != nullptr && (a->ld_ctx % 8) == 0, MERLOT_EINVAL, "attention_fwd: ctx missing or ld_ctx %% 8 != 0");
  AttnDev p; fill_dev(a, &p);
  CUtensorMap tq, tkv;
  rc = make_tmap_bf16_2d(&tq, a->qkv, (uint64_t)a->ld_qkv, (uint64_t)a->B * a->S, (uint64_t)a->ld_qkv, 64, AT_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tkv, a->qkv, (uint64_t)a->ld_qkv, (uint64_t)a->B * a->S, (uint64_t)a->ld_qkv, 64, FK);
  if (rc) return rc;
  MB_REQUIRE(a->S <= MAX_MASK_WORDS * 32 - AT_N, MERLOT_ESHAPE, "attention_fwd: sequence longer than %d keys", MAX_MASK_WORDS * 32 - AT_N);
  static bool attr = false;
  if (!attr) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr = true;
  }
  dim3 grid(ceil_div(a->S, AT_M), a->heads, a->B);
  if (a->valid) MB_CHECK_CUDA(launch_pdl(attn_fwd_kernel<true>, grid, dim3(128), FWD_SMEM, stream, tq, tkv, p));
  else MB_CHECK_CUDA(launch_pdl(attn_fwd_kernel<false>, grid, dim3(128), FWD_SMEM, stream, tq, tkv, p));
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

extern "C" int merlot_attention_bwd_dq_parts(int S) {
  const int n_kv = ceil_div(S, AT_N);
  return n_kv <= MAX_DQ_PARTS ? n_kv : 0;
}

extern "C" size_t merlot_attention_bwd_workspace_bytes(int B, int S, int heads) {
  const int parts = merlot_attention_bwd_dq_parts(S);
  return (size_t)(parts > 0 ? parts : 1) * (size_t)B * S * heads * 64 * sizeof(float);
}

extern "C" int merlot_attention_bwd(const merlot_attn_t* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_common(a);
  if (rc) return rc;
  MB_REQUIRE(a->ctx && a->d_ctx && a->lse && a->dsum && a->dq_accum && a->dqkv, MERLOT_EINVAL,
             "attention_bwd: ctx, d_ctx, lse, dsum, dq_accum and dqkv are all required");
  MB_REQUIRE((a->ld_ctx % 8) == 0 && (a->ld_dqkv % 8) == 0 && (a->ld_dq % 4) == 0, MERLOT_ESHAPE,
             "attention_bwd: leading dimensions must keep 16-byte alignment");
  AttnDev p; fill_dev(a, &p);
  const int H = p.H;
  const long long tokens = (long long)a->B * a->S;
  const int parts = merlot_attention_bwd_dq_parts(a->S);
  p.dq_part_stride = (size_t)tokens * a->ld_dq;
  {  // D = rowsum(dO * O)
    const long long threads = tokens * a->heads * 8;
    MB_CHECK_CUDA(launch_pdl(attn_dsum_kernel, dim3((unsigned)ceil_div_ll(threads, 256)), dim3(256), 0, stream,
                             reinterpret_cast<const bf16*>(a->ctx), reinterpret_cast<const bf16*>(a->d_ctx), a->ld_ctx, a->dsum,
                             a->B, a->S, a->heads));
    MB_CHECK_LAUNCH();
  }
  CUtensorMap tkv, tq, tdo;
  rc = make_tmap_bf16_2d(&tkv, a->qkv, (uint64_t)a->ld_qkv, (uint64_t)tokens, (uint64_t)a->ld_qkv, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tq, a->qkv, (uint64_t)a->ld_qkv, (uint64_t)tokens, (uint64_t)a->ld_qkv, 64, BQ);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tdo, a->d_ctx, (uint64_t)a->ld_ctx, (uint64_t)tokens, (uint64_t)a->ld_ctx, 64, BQ);
  if (rc) return rc;
  MB_REQUIRE(a->S <= MAX_MASK_WORDS * 32 - AT_M, MERLOT_ESHAPE, "attention_bwd: sequence longer than %d keys", MAX_MASK_WORDS * 32 - AT_M);
  static bool attr = false;
  if (!attr) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr = true;
  }
  const long long items = (long long)ceil_div(a->S, AT_N) * a->heads * a->B;
  dim3 grid((unsigned)(items < num_sms() ? items : num_sms()));  // persistent: one CTA per SM walks the (key tile, head, batch) items
  if (parts > 0) {
    if (a->valid) MB_CHECK_CUDA(launch_pdl(attn_bwd_kernel<true, false>, grid, dim3(BWD_THREADS), BWD_SMEM, stream, tkv, tq, tdo, p));
    else MB_CHECK_CUDA(launch_pdl(attn_bwd_kernel<false, false>, grid, dim3(BWD_THREADS), BWD_SMEM, stream, tkv, tq, tdo, p));
  } else {
    if (a->valid) MB_CHECK_CUDA(launch_pdl(attn_bwd_kernel<true, true>, grid, dim3(BWD_THREADS), BWD_SMEM, stream, tkv, tq, tdo, p));
    else MB_CHECK_CUDA(launch_pdl(attn_bwd_kernel<false, true>, grid, dim3(BWD_THREADS), BWD_SMEM, stream, tkv, tq, tdo, p));
  }
  MB_CHECK_LAUNCH();
  {
    long long slabs = ceil_div_ll(tokens, 64);
    if (slabs > 128) slabs = 128;
    MB_CHECK_CUDA(launch_pdl(attn_dqkv_finish_kernel, dim3(ceil_div(3 * H, 256), (unsigned)slabs), dim3(256), 0, stream, a->dq_accum,
                             a->ld_dq, p.dqkv, a->ld_dqkv, tokens, H, a->d_bias_qkv, parts, p.dq_part_stride, a->scale));
    MB_CHECK_LAUNCH();
  }
  return MERLOT_OK;
}

// attention_log (model/modeling.py:186-203): 4 normalised block sums from the split column sums.
//   c_viz[b,k] / c_lang[b,k] = sum over (layers, valid queries in the viz / lang piece) of head-mean probabilities
//   out = {lang2lang, lang2viz, viz2lang, viz2viz}  (names are `from`2`to`: keys are `from`, queries are `to`), sum = 1
__global__ void attn_log_blocks_kernel(const float* __restrict__ c_viz, const float* __restrict__ c_lang, const uint8_t* __restrict__ valid,
                                       int B, int S, int P, float* __restrict__ out) {
  __shared__ float red[4][256];
  float a[4] = {0.f, 0.f, 0.f, 0.f};  // [to_viz_from_viz, to_viz_from_lang, to_lang_from_viz, to_lang_from_lang]
  for (int i = threadIdx.x; i < B * S; i += 256) {
    const int k = i % S;
    const float vk = valid[i] ? 1.f : 0.f;
    const int from_lang = k >= P;
    a[0 + from_lang] += c_viz[i] * vk;
    a[2 + from_lang] += c_lang[i] * vk;
  }
  for (int j = 0; j < 4; ++j) red[j][threadIdx.x] = a[j];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float tot = red[0][0] + red[1][0] + red[2][0] + red[3][0];
    out[0] = red[3][0] / tot;  // lang2lang : keys lang, queries lang
    out[1] = red[1][0] / tot;  // lang2viz  : keys lang, queries viz
    out[2] = red[2][0] / tot;  // viz2lang  : keys viz,  queries lang
    out[3] = red[0][0] / tot;  // viz2viz
  }
}

extern "C" int merlot_attention_log_blocks(const float* c_viz, const float* c_lang, const void* valid, int B, int S, int P, float* out4,
                                           void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(c_viz && c_lang && valid && out4, MERLOT_EINVAL, "attention_log_blocks: null pointer");
  attn_log_blocks_kernel<<<1, 256, 0, stream>>>(c_viz, c_lang, reinterpret_cast<const uint8_t*>(valid), B, S, P, out4);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

extern "C" int merlot_attention_colsum(const merlot_attn_t* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = check_common(a);
  if (rc) return rc;
  MB_REQUIRE(a->lse && a->colsum, MERLOT_EINVAL, "attention_colsum: lse and colsum are required");
  AttnDev p; fill_dev(a, &p);
  CUtensorMap tm;
  rc = make_tmap_bf16_2d(&tm, a->qkv, (uint64_t)a->ld_qkv, (uint64_t)a->B * a->S, (uint64_t)a->ld_qkv, 64, 128);
  if (rc) return rc;
  MB_REQUIRE(a->S <= MAX_MASK_WORDS * 32 - AT_M, MERLOT_ESHAPE, "attention_colsum: sequence longer than %d keys", MAX_MASK_WORDS * 32 - AT_M);
  static bool attr = false;
  if (!attr) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_colsum_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
    MB_CHECK_CUDA(cudaFuncSetAttribute(attn_colsum_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CS_SMEM));
    attr = true;
  }
  dim3 grid(ceil_div(a->S, AT_N), a->heads, a->B);
  if (a->valid) MB_CHECK_CUDA(launch_pdl(attn_colsum_kernel<true>, grid, dim3(128), CS_SMEM, stream, tm, p));
  else MB_CHECK_CUDA(launch_pdl(attn_colsum_kernel<false>, grid, dim3(128), CS_SMEM, stream, tm, p));
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}
