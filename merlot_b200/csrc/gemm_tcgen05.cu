// K1: persistent, warp-specialised bf16 GEMM for sm_100a.
//   warp 0      : TMA producer (one elected lane) -- A/B tiles -> 128B-swizzled smem ring
//   warp 1      : TMEM allocator + MMA issuer (one elected lane) -- tcgen05.mma 128 x BN x 16, fp32 accum in TMEM
//   warps 2..5  : epilogue -- tcgen05.ld TMEM -> registers -> fused epilogue -> vectorised global stores
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue; 2 accumulator stages so
// the epilogue of tile i overlaps the main loop of tile i+1), and a static persistent tile schedule.
//
// Replaces the reference's tf.layers.dense / tf.matmul call sites listed in include/merlot_b200.h (K1).
#include "gemm_common.cuh"

#include <stdlib.h>

#include <vector>

namespace mb {

// TS = true: bf16 outputs leave through 128B-swizzled smem staging boxes and TMA stores (fully coalesced, edge clipping
// by the tensor map).  TS = false: direct register->global stores (fp32 outputs / atomics).
template <int BN, bool A_MN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_o1, const __grid_constant__ CUtensorMap tma_o2, const GemmDev p) {
  using Cfg = GemmCfg<BN, EPI>;
  constexpr bool TS = EPI != 0;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms need 1024-byte alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + (TS ? STAGING_BYTES : 0));
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (TS) { tma_prefetch_desc(&tma_o1); tma_prefetch_desc(&tma_o2); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_WARPS);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();  // everything above overlapped the previous kernel's tail

  const int num_tiles = p.m_blocks * p.n_blocks * p.splits;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile % p.splits;
        const int mn = tile / p.splits;
        const int n_blk = mn % p.n_blocks;
        const int m_blk = mn / p.n_blocks;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c)
              tma_load_2d(sa + c * (BLOCK_K * 128), &tma_a, &full_bar[stage], m_blk * BLOCK_M + c * 64, kb * BLOCK_K);
          } else {
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * (BLOCK_K * 128), &tma_b, &full_bar[stage], n_blk * BN + c * 64, kb * BLOCK_K);
          } else {
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BLOCK_K, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int split = tile % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t da = A_MN ? desc_mnmajor(a_addr, k, BLOCK_K * 128) : desc_kmajor(a_addr, k);
            const uint64_t db = B_MN ? desc_mnmajor(b_addr, k, BLOCK_K * 128) : desc_kmajor(b_addr, k);
            umma_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 2+EPI_WARPS) =====================
    const int e = warp - 2;
    const int quad = warp & 3;          // TMEM lane window this warp may touch: lanes [32*quad, 32*quad+32)
    const int half = e >> 2;            // which half of the tile's columns this warp owns
    const int row_in_tile = quad * 32 + lane;
    const bool issuer = (e == 0 && lane == 0);
    const bool dual = (p.flags & MERLOT_GEMM_GELU) && p.out2 != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mn = tile / p.splits;
      const int n_blk = mn % p.n_blocks;
      const int m_blk = mn / p.n_blocks;
      const int row = m_blk * BLOCK_M + row_in_tile;
      const bool in_range = row < p.M;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
      if (!TS) {
        constexpr int CH = BN / 64;  // 32-column chunks per warp
#pragma unroll 1
        for (int c = 0; c < CH; ++c) {
          uint32_t r[32];
          const int tcol = half * (BN / 2) + c * 32;
          tmem_ld_32x32(taddr + tcol, r);
          tmem_wait_ld();
          const int col0 = n_blk * BN + tcol;
          if (in_range && col0 < p.N) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = col0 + g * 8;
              if (col < p.N) {
                float v[8], pre[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
                epi_math8(p, row, col, true, v, pre);
                epi_store_direct(p, row, col, v, pre);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      } else {
        // Staged epilogue in phases of PCOLS accumulator columns; each phase fills the two 16 KB staging boxes
        // ([128 rows][128 B], 128B-swizzled) and hands them to TMA:
        //   bf16 single output : PCOLS = 128 -> box h = columns [64h, 64h+64) of the phase            (TMA store)
        //   bf16 pre+act output: PCOLS = 64  -> box 0 = pre-activation, box 1 = gelu, same 64 columns  (2 TMA stores)
        //   fp32 split-K accum : PCOLS = 64  -> box h = 32 fp32 columns                                (TMA reduce-add)
        const int pcols_max = (EPI == 2 || dual) ? 64 : 128;
        const int phases = (BN + pcols_max - 1) / pcols_max;
        for (int ph = 0; ph < phases; ++ph) {
          if (issuer) tma_store_wait_read_all();  // staging is free again
          named_bar_sync(1, 32 * EPI_WARPS);
          const int pcols = min(pcols_max, BN - ph * pcols_max);
          const int wcols = pcols >> 1;           // columns per warp half: 64 or 32
          const int tcol0 = ph * pcols_max + half * wcols;
#pragma unroll 1
          for (int c = 0; c < wcols / 32; ++c) {
            uint32_t r[32];
            const int tcol = tcol0 + c * 32;
            tmem_ld_32x32(taddr + tcol, r);
            tmem_wait_ld();
            const int col0 = n_blk * BN + tcol;
            if (EPI == 2) {
              uint8_t* box = staging + half * 16384;
#pragma unroll
              for (int g = 0; g < 8; ++g)
                *reinterpret_cast<uint4*>(box + sw128_offset((uint32_t)row_in_tile, (uint32_t)g)) =
                    make_uint4(__float_as_uint(__uint_as_float(r[g * 4 + 0]) * p.alpha), __float_as_uint(__uint_as_float(r[g * 4 + 1]) * p.alpha),
                               __float_as_uint(__uint_as_float(r[g * 4 + 2]) * p.alpha), __float_as_uint(__uint_as_float(r[g * 4 + 3]) * p.alpha));
            } else {
              const int pcol = half * wcols + c * 32;  // column inside the phase
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int col = col0 + g * 8;
                float v[8], pre[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
                if (col < p.N) epi_math8(p, row, col, in_range, v, pre);
                const int chunk16 = ((pcol & 63) >> 3) + g;
                if (dual) {
                  stage_bf16x8(staging, row_in_tile, chunk16, pre);
                  stage_bf16x8(staging + 16384, row_in_tile, chunk16, v);
                } else {
                  stage_bf16x8(staging + (pcol >> 6) * 16384, row_in_tile, chunk16, v);
                }
              }
            }
          }
          if (ph == phases - 1) {  // all TMEM reads of this accumulator stage are done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 32 * EPI_WARPS);
          if (issuer) {
            const int r0 = m_blk * BLOCK_M;
            const int c0 = n_blk * BN + ph * pcols_max;
            if (EPI == 2) {
              for (int b2 = 0; b2 < 2; ++b2)
                if (b2 * 32 < pcols && c0 + b2 * 32 < p.N) tma_reduce_add_2d(&tma_o1, staging + b2 * 16384, c0 + b2 * 32, r0);
            } else if (dual) {
              if (c0 < p.N) {
                tma_store_2d(&tma_o1, staging, c0, r0);
                tma_store_2d(&tma_o2, staging + 16384, c0, r0);
              }
            } else {
              for (int b2 = 0; b2 < 2; ++b2)
                if (b2 * 64 < pcols && c0 + b2 * 64 < p.N) tma_store_2d(&tma_o1, staging + b2 * 16384, c0 + b2 * 64, r0);
            }
            tma_store_commit();
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (TS && issuer) tma_store_wait_all();  // global writes complete before the CTA exits
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// Optional per-launch timing (bench.py roofline): CUDA events on the launching stream around every GEMM launch.
struct ProfRec { cudaEvent_t e0, e1; double flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;

void* gemm_prof_before(double flops, cudaStream_t stream) {
  if (!g_prof_on) return nullptr;
  ProfRec* rec = new ProfRec;
  rec->flops = flops;
  cudaEventCreate(&rec->e0);
  cudaEventCreate(&rec->e1);
  cudaEventRecord(rec->e0, stream);
  return rec;
}
void gemm_prof_after(void* tok, cudaStream_t stream) {
  if (!tok) return;
  ProfRec* rec = reinterpret_cast<ProfRec*>(tok);
  cudaEventRecord(rec->e1, stream);
  g_prof.push_back(*rec);
  delete rec;
}

template <int BN, bool A_MN, bool B_MN, int EPI>
static int launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to1, const CUtensorMap& to2,
                            const GemmDev& p, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, EPI>;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, EPI>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_TOTAL));
    attr_set = true;
  }
  void* tok = gemm_prof_before(2.0 * (double)p.M * (double)p.N * (double)p.K, stream);
  MB_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_TOTAL, stream, ta, tb, to1, to2, p));
  MB_CHECK_LAUNCH();
  gemm_prof_after(tok, stream);
  return MERLOT_OK;
}

}  // namespace mb

namespace mb {
int launch_gemm_pair(bool a_mn, bool b_mn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to1,
                     const CUtensorMap& to2, const GemmDev& p, int grid, cudaStream_t stream);
}

using namespace mb;

extern "C" int merlot_gemm_bf16(const merlot_gemm_t* g, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(g != nullptr, MERLOT_EINVAL, "gemm: null descriptor");
  MB_REQUIRE(g->a && g->b && g->out, MERLOT_EINVAL, "gemm: null operand pointer");
  MB_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, MERLOT_ESHAPE, "gemm: non-positive dims M=%d N=%d K=%d", g->M, g->N,
             g->K);
  MB_REQUIRE((g->lda % 8) == 0 && (g->ldb % 8) == 0, MERLOT_ESHAPE,
             "gemm: lda/ldb must be multiples of 8 bf16 elements for TMA (lda=%d ldb=%d)", g->lda, g->ldb);
  MB_REQUIRE(((uintptr_t)g->a % 16) == 0 && ((uintptr_t)g->b % 16) == 0, MERLOT_ESHAPE,
             "gemm: operand base pointers must be 16-byte aligned");
  MB_REQUIRE(g->lda >= (g->a_mn_major ? g->M : g->K) && g->ldb >= (g->b_mn_major ? g->N : g->K), MERLOT_ESHAPE,
             "gemm: leading dimension smaller than the row length");
  const bool out_f32 = g->flags & MERLOT_GEMM_OUT_F32;
  MB_REQUIRE(!(g->flags & MERLOT_GEMM_ATOMIC) || out_f32, MERLOT_EINVAL, "gemm: ATOMIC requires OUT_F32");
  MB_REQUIRE(!(g->flags & MERLOT_GEMM_MUL_DGELU) || g->aux, MERLOT_EINVAL, "gemm: MUL_DGELU requires aux");
  MB_REQUIRE(g->ld_out >= g->N, MERLOT_ESHAPE, "gemm: ld_out < N");
  if (!out_f32)
    MB_REQUIRE((g->ld_out % 8) == 0 && ((uintptr_t)g->out % 16) == 0, MERLOT_ESHAPE,
               "gemm: bf16 output needs ld_out %% 8 == 0 and a 16-byte aligned base");
  if (g->resid) MB_REQUIRE((g->ld_resid % 8) == 0, MERLOT_ESHAPE, "gemm: ld_resid %% 8 != 0");
  if (g->aux) MB_REQUIRE((g->ld_aux % 8) == 0, MERLOT_ESHAPE, "gemm: ld_aux %% 8 != 0");
  if (g->out2) MB_REQUIRE((g->ld_out2 % 8) == 0, MERLOT_ESHAPE, "gemm: ld_out2 %% 8 != 0");
  if (g->flags & MERLOT_GEMM_DROPOUT)
    MB_REQUIRE((g->N % 8) == 0 && g->dropout_p >= 0.f && g->dropout_p < 1.f, MERLOT_ESHAPE,
               "gemm: dropout needs N %% 8 == 0 and 0 <= p < 1");

  GemmDev p;
  memset(&p, 0, sizeof(p));
  p.M = g->M; p.N = g->N; p.K = g->K;
  p.out = g->out; p.ld_out = g->ld_out; p.out2 = g->out2; p.ld_out2 = g->ld_out2;
  p.bias = g->bias;
  p.resid = reinterpret_cast<const bf16*>(g->resid); p.ld_resid = g->ld_resid;
  p.aux = reinterpret_cast<const bf16*>(g->aux); p.ld_aux = g->ld_aux;
  p.alpha = g->alpha;
  p.flags = g->flags;
  if ((g->flags & MERLOT_GEMM_DROPOUT) && g->dropout_p > 0.f) {
    p.drop_thresh16 = (uint32_t)(g->dropout_p * 65536.0f + 0.5f);
    p.drop_scale = 1.0f / (1.0f - g->dropout_p);
    p.seed = g->dropout_seed;
    p.site = g->dropout_site;
  } else {
    p.flags &= ~MERLOT_GEMM_DROPOUT;
  }

  const int sms = num_sms();
  p.m_blocks = ceil_div(g->M, BLOCK_M);
  p.num_kb = ceil_div(g->K, BLOCK_K);
  // ---- tile width: minimise wave-quantisation loss; BN=256 halves per-FLOP smem traffic so it wins ties ----
  int bn = g->block_n;
  // CTA-pair kernel (cta_group::2, 256 x 256 tile): block_n = -256 forces it; MERLOT_GEMM_PAIR=1 lets the heuristic pick it
  // (measured: 8192^3 1464 vs 1305 TF/s, split-K wgrad 1281 vs 1208; a wash at K = 768 where tile quantisation dominates)
  static const int pair_auto = [] { const char* e = getenv("MERLOT_GEMM_PAIR"); return e ? atoi(e) : 1; }();
  bool pair = false;
  if (bn == -256) { pair = true; bn = 256; }
  if (bn == 0 && (g->flags & MERLOT_GEMM_ATOMIC)) bn = 256;  // wgrad: split-K fills the machine, wide tiles halve smem traffic
  if (bn == 0) {
    const bool dual = (g->flags & MERLOT_GEMM_GELU) && g->out2 != nullptr;
    double best = -1;
    for (int cand : {256, 192, 128}) {
      if (dual && cand == 192) continue;  // pre+act staging works in 128-column phases
      long long tiles = (long long)p.m_blocks * ceil_div(g->N, cand);
      long long waves = ceil_div_ll(tiles, sms);
      double useful = (double)g->N / (double)(ceil_div(g->N, cand) * cand);
      double eff = (double)tiles / (double)(waves * sms) * useful * (cand == 256 ? 1.0 : (cand == 192 ? 0.97 : 0.80));
      if (eff > best + 1e-9) { best = eff; bn = cand; }
    }
  }
  MB_REQUIRE(bn == 128 || bn == 192 || bn == 256, MERLOT_EINVAL, "gemm: block_n must be 0, 128, 192 or 256 (got %d)", bn);
  MB_REQUIRE(!(bn == 192 && (g->flags & MERLOT_GEMM_GELU) && g->out2), MERLOT_EINVAL, "gemm: block_n 192 cannot be used with a dual (pre+act) output");
  p.n_blocks = ceil_div(g->N, bn);
  if (!pair && pair_auto == 2 && bn == 256 && g->M > 256) pair = true;  // 2 = everywhere (experiments)
  if (!pair && pair_auto == 1 && bn == 256 && g->M > 256 && ((g->flags & MERLOT_GEMM_ATOMIC) || g->K >= 4096)) pair = true;
  const int units = pair ? sms / 2 : sms;                                   // schedulable CTAs or CTA pairs
  const int m_tiles = pair ? ceil_div(g->M, 256) : p.m_blocks;
  // ---- split-K (wgrad): fill the machine when the MN tile count is small ----
  int splits = g->splits;
  const int mn_tiles = m_tiles * p.n_blocks;
  if (splits <= 0) {
    splits = 1;
    if ((g->flags & MERLOT_GEMM_ATOMIC) && mn_tiles < units) {
      splits = units / mn_tiles;
      int max_by_k = p.num_kb / 4 > 0 ? p.num_kb / 4 : 1;  // keep >= 4 k-blocks per split
      if (splits > max_by_k) splits = max_by_k;
      if (splits < 1) splits = 1;
    }
  }
  MB_REQUIRE(splits == 1 || (g->flags & MERLOT_GEMM_ATOMIC), MERLOT_EINVAL, "gemm: splits>1 requires ATOMIC");
  if (splits > p.num_kb) splits = p.num_kb;
  p.kb_per_split = ceil_div(p.num_kb, splits);
  p.splits = ceil_div(p.num_kb, p.kb_per_split);  // no empty splits

  CUtensorMap ta, tb;
  int rc;
  if (g->a_mn_major)
    rc = make_tmap_bf16_2d(&ta, g->a, (uint64_t)g->M, (uint64_t)g->K, (uint64_t)g->lda, 64, BLOCK_K);
  else
    rc = make_tmap_bf16_2d(&ta, g->a, (uint64_t)g->K, (uint64_t)g->M, (uint64_t)g->lda, BLOCK_K, BLOCK_M);
  if (rc) return rc;
  if (g->b_mn_major)
    rc = make_tmap_bf16_2d(&tb, g->b, (uint64_t)g->N, (uint64_t)g->K, (uint64_t)g->ldb, 64, BLOCK_K);
  else
    rc = make_tmap_bf16_2d(&tb, g->b, (uint64_t)g->K, (uint64_t)g->N, (uint64_t)g->ldb, BLOCK_K, (uint32_t)(pair ? bn / 2 : bn));
  if (rc) return rc;

  const long long tiles = (long long)mn_tiles * p.splits;
  const int grid = pair ? 2 * (int)(tiles < units ? tiles : units) : (int)(tiles < sms ? tiles : sms);

  // epilogue route: 1 = bf16 tiles through swizzled smem staging + TMA store; 2 = fp32 split-K accumulation through TMA
  // reduce-add (plain alpha*acc only); 0 = direct register->global stores (other fp32 outputs)
  const bool plain = !g->bias && !g->resid && !(g->flags & (MERLOT_GEMM_GELU | MERLOT_GEMM_MUL_DGELU | MERLOT_GEMM_DROPOUT));
  const int epi = !out_f32 ? 1
                  : ((g->flags & MERLOT_GEMM_ATOMIC) && plain && (g->ld_out % 4) == 0 && ((uintptr_t)g->out % 16) == 0) ? 2 : 0;
  CUtensorMap to1, to2;
  memset(&to1, 0, sizeof(to1));
  memset(&to2, 0, sizeof(to2));
  if (epi == 1) {
    const bool dual = (g->flags & MERLOT_GEMM_GELU) && g->out2 != nullptr;
    if (dual) MB_REQUIRE(((uintptr_t)g->out2 % 16) == 0, MERLOT_ESHAPE, "gemm: out2 must be 16-byte aligned");
    rc = make_tmap_bf16_2d(&to1, g->out, (uint64_t)g->N, (uint64_t)g->M, (uint64_t)g->ld_out, 64, BLOCK_M);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&to2, dual ? g->out2 : g->out, (uint64_t)g->N, (uint64_t)g->M,
                           (uint64_t)(dual ? g->ld_out2 : g->ld_out), 64, BLOCK_M);
    if (rc) return rc;
  } else if (epi == 2) {
    rc = make_tmap_f32_2d(&to1, g->out, (uint64_t)g->N, (uint64_t)g->M, (uint64_t)g->ld_out, 32, BLOCK_M);
    if (rc) return rc;
  }
  if (pair) return launch_gemm_pair(g->a_mn_major != 0, g->b_mn_major != 0, epi, ta, tb, to1, to2, p, grid, stream);
#define MB_GEMM_DISPATCH2(BN_, EPI_)                                                                                  \
  if (g->a_mn_major && g->b_mn_major) return launch_gemm_inst<BN_, true, true, EPI_>(ta, tb, to1, to2, p, grid, stream);   \
  if (!g->a_mn_major && g->b_mn_major) return launch_gemm_inst<BN_, false, true, EPI_>(ta, tb, to1, to2, p, grid, stream); \
  if (!g->a_mn_major && !g->b_mn_major) return launch_gemm_inst<BN_, false, false, EPI_>(ta, tb, to1, to2, p, grid, stream); \
  return launch_gemm_inst<BN_, true, false, EPI_>(ta, tb, to1, to2, p, grid, stream);
#define MB_GEMM_DISPATCH(BN_)                  \
  if (epi == 1) { MB_GEMM_DISPATCH2(BN_, 1) }  \
  if (epi == 2) { MB_GEMM_DISPATCH2(BN_, 2) }  \
  MB_GEMM_DISPATCH2(BN_, 0)
  if (bn == 256) { MB_GEMM_DISPATCH(256) }
  if (bn == 192) { MB_GEMM_DISPATCH(192) }
  MB_GEMM_DISPATCH(128)
#undef MB_GEMM_DISPATCH
#undef MB_GEMM_DISPATCH2
}

extern "C" void merlot_gemm_profile_begin(void) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on = true;
}

extern "C" int merlot_gemm_profile_end(double* total_ms, double* total_flops, long long* launches) {
  g_prof_on = false;
  MB_CHECK_CUDA(cudaDeviceSynchronize());
  double ms = 0, fl = 0;
  for (auto& r : g_prof) {
    float t = 0.f;
    MB_CHECK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t; fl += r.flops;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = (long long)g_prof.size();
  g_prof.clear();
  return MERLOT_OK;
}
