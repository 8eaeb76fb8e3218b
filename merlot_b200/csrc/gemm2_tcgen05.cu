// K1, CTA-pair variant: two CTAs of a cluster (one per SM of a TPC) cooperate on one 256 x 256 output tile with
// tcgen05.mma.cta_group::2.  Each CTA TMA-loads its own 128 A rows and HALF of the B tile (128 of the 256 N rows) into its
// own shared memory, so per-SM shared-memory fill and L2->SM traffic per FLOP drop by a third (32 KB instead of 48 KB per
// k-block) and the smem ring is 6 stages deep instead of 4 -- the 1-CTA kernel is TMA-latency bound at K-block = 64.
//   both CTAs : warp 0 = TMA producer (arms / credits the LEADER's full barrier), warps 2..9 = epilogue of their 128 rows
//   leader    : warp 1 lane 0 issues the MMAs for the pair; tcgen05.commit multicasts to both CTAs' barriers
// Everything else (fused epilogues, TMA-store / reduce-add staging, tile schedule) is shared with gemm_tcgen05.cu.
#include "gemm_common.cuh"

namespace mb {

constexpr int A2_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB: this CTA's 128 rows of A
template <int BN>  // pair tile width: 256, or 192 (K-major B only) for N = 768-class outputs whose 256-wide tiling wastes a wave
struct Pair {
  static constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;  // 16 / 12 KB: this CTA's half of B
  static constexpr int STAGE_BYTES = A2_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_LIMIT - STAGING_BYTES) / STAGE_BYTES;  // 6 / 6
  static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 256 + BIAS_SLOT_BYTES;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to the barrier at `bar_cluster_addr` (the pair leader's barrier)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the pair's MMAs issued so far have completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_multicast(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int BN, bool A_MN, bool B_MN, int EPI, int FL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const GemmDev p) {
  static_assert(BN == 256 || (BN == 192 && !B_MN), "an MN-major B half must be a whole number of 64-column TMA boxes");
  constexpr int STAGES2 = Pair<BN>::STAGES, B2_BYTES = Pair<BN>::B_BYTES, STAGE2_BYTES = Pair<BN>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES2 * A2_BYTES;
  uint8_t* staging = smem + STAGES2 * STAGE2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* full_bar = bars;                       // used in the leader only
  uint64_t* empty_bar = bars + STAGES2;            // per CTA, signalled by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * STAGES2;        // per CTA, multicast commit
  uint64_t* tmem_empty = bars + 2 * STAGES2 + 2;   // leader only: both CTAs' epilogue warps arrive
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);
  float* bias_slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform (see gemm_kernel.cuh)
  const int lane = threadIdx.x & 31;
  const bool elected = elect_one();                                        // the single issuing lane of a warp
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES2; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_ptr_smem, 512);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr_smem, 0);
  pdl_wait();

  const int m2_blocks = (p.M + 255) / 256;  // pair tiles along M
  const int num_tiles = m2_blocks * p.n_blocks * p.splits;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elected) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int split = tile % p.splits;
        const int mn = tile / p.splits;
        const int n_blk = mn % p.n_blocks;
        const int m2 = mn / p.n_blocks;
        const int m0 = m2 * 256 + (int)rank * BLOCK_M;      // this CTA's A rows
        const int n0 = n_blk * BN + (int)rank * (BN / 2);   // this CTA's half of B
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          const uint32_t fb = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;  // same offset in the pair's leader (even) CTA
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);
          uint8_t* sa = smem_a + stage * A2_BYTES;
          uint8_t* sb = smem_b + stage * B2_BYTES;
          if (A_MN) {
#pragma unroll
            for (int c = 0; c < BLOCK_M / 64; ++c) tma_load_2d_pair(sa + c * (BLOCK_K * 128), &tma_a, fb, m0 + c * 64, kb * BLOCK_K);
          } else {
            tma_load_2d_pair(sa, &tma_a, fb, kb * BLOCK_K, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int c = 0; c < (BN / 2) / 64; ++c) tma_load_2d_pair(sb + c * (BLOCK_K * 128), &tma_b, fb, n0 + c * 64, kb * BLOCK_K);
          } else {
            tma_load_2d_pair(sb, &tma_b, fb, kb * BLOCK_K, n0);
          }
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) if (elected) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int split = tile % p.splits;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * A2_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * B2_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / 16; ++k) {
            const uint64_t da = A_MN ? desc_mnmajor(a_addr, k, BLOCK_K * 128) : desc_kmajor(a_addr, k);
            const uint64_t db = B_MN ? desc_mnmajor(b_addr, k, BLOCK_K * 128) : desc_kmajor(b_addr, k);
            umma2_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma2_commit_multicast(&empty_bar[stage]);  // frees this smem slot in both CTAs
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma2_commit_multicast(&tmem_full[acc]);  // accumulators complete in both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    const int e = warp - 2;
    const int quad = warp & 3;
    const int half = e >> 2;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t leader_tmem_empty0 = mapa_u32(smem_u32(&tmem_empty[0]), 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    const int rtf = FL == F_GENERIC ? epi_features(p) : FL;
    EpiCarry cy;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int mn = tile / p.splits;
      const int n_blk = mn % p.n_blocks;
      const int m2 = mn / p.n_blocks;
      const int row0 = m2 * 256 + (int)rank * BLOCK_M;
      const int row = row0 + row_in_tile;
      const bool in_range = row < p.M;
      if (EPI != 0) epilogue_prefetch<BN, EPI, FL>(p, rtf, row0 + quad * 32, n_blk * BN, half, lane, cy);  // hidden by the wait
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
      if (EPI == 0) {
        constexpr int CH = BN / 64;
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
        if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + acc * 8);
      } else {
        epilogue_tile_loop<BN, EPI, FL>(p, rtf, staging + e * 4096, bias_slots + e * 32, taddr, row0 + quad * 32, n_blk * BN, half, lane, cy, [&] {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_tmem_empty0 + acc * 8);
        });
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI, int FL>
static int launch_gemm2_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int grid, cudaStream_t stream) {
  auto kern = gemm2_bf16_kernel<BN, A_MN, B_MN, EPI, FL>;
  constexpr int SMEM2_TOTAL = Pair<BN>::SMEM_TOTAL;
  static bool attr_set = false;
  if (!attr_set) {
    MB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_TOTAL));
    attr_set = true;
  }
  void* tok = gemm_prof_before(2.0 * (double)p.M * (double)p.N * (double)p.K, stream);
  MB_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), SMEM2_TOTAL, stream, ta, tb, p));
  MB_CHECK_LAUNCH();
  gemm_prof_after(tok, stream);
  return MERLOT_OK;
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm2_mn(int epi, int fl, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int grid, cudaStream_t stream) {
  if (epi == 0) return launch_gemm2_inst<BN, A_MN, B_MN, 0, F_GENERIC>(ta, tb, p, grid, stream);
  if (epi == 2) return launch_gemm2_inst<BN, A_MN, B_MN, 2, F_ALPHA>(ta, tb, p, grid, stream);
  switch (fl) {  // same specialised feature sets as the 1-CTA kernel (gemm_kernel.cuh)
    case 0: return launch_gemm2_inst<BN, A_MN, B_MN, 1, 0>(ta, tb, p, grid, stream);
    case F_BIAS: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_BIAS>(ta, tb, p, grid, stream);
    case F_BIAS | F_GELU | F_DUAL: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_BIAS | F_GELU | F_DUAL>(ta, tb, p, grid, stream);
    case F_DGELU: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_DGELU>(ta, tb, p, grid, stream);
    case F_BIAS | F_GELU | F_DUAL | F_GRADOUT: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_BIAS | F_GELU | F_DUAL | F_GRADOUT>(ta, tb, p, grid, stream);
    case F_MULAUX: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_MULAUX>(ta, tb, p, grid, stream);
    case F_BIAS | F_RESID: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_BIAS | F_RESID>(ta, tb, p, grid, stream);
    case F_BIAS | F_RESID | F_DROP: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_BIAS | F_RESID | F_DROP>(ta, tb, p, grid, stream);
    case F_RESID: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_RESID>(ta, tb, p, grid, stream);
    default: return launch_gemm2_inst<BN, A_MN, B_MN, 1, F_GENERIC>(ta, tb, p, grid, stream);
  }
}

// Called by merlot_gemm_bf16 when the pair kernel is selected.  `p` arrives with n_blocks for the tile width `bn`.
int launch_gemm_pair(int bn, bool a_mn, bool b_mn, int epi, int fl, const CUtensorMap& ta, const CUtensorMap& tb, const GemmDev& p, int grid,
                     cudaStream_t stream) {
  if (bn == 192) {
    if (b_mn) return set_error(MERLOT_EINVAL, "gemm: the 192-wide pair tile needs a K-major B operand");
    if (a_mn) return launch_gemm2_mn<192, true, false>(epi, fl, ta, tb, p, grid, stream);
    return launch_gemm2_mn<192, false, false>(epi, fl, ta, tb, p, grid, stream);
  }
  if (a_mn && b_mn) return launch_gemm2_mn<256, true, true>(epi, fl, ta, tb, p, grid, stream);
  if (!a_mn && b_mn) return launch_gemm2_mn<256, false, true>(epi, fl, ta, tb, p, grid, stream);
  if (!a_mn && !b_mn) return launch_gemm2_mn<256, false, false>(epi, fl, ta, tb, p, grid, stream);
  return launch_gemm2_mn<256, true, false>(epi, fl, ta, tb, p, grid, stream);
}

}  // namespace mb
