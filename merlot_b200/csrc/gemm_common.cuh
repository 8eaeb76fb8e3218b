// Shared pieces of the K1 GEMM kernels (1-CTA and CTA-pair variants): tile constants, the device-side problem descriptor and the
// fused epilogue math / store helpers.
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace mb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle atom
constexpr int EPI_WARPS = 8;  // two warps per TMEM lane quadrant, each owning half of the tile's columns
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int BIAS_SLOT_BYTES = 1024;                       // 8 epilogue warps x 32 fp32 bias values of the current chunk
constexpr int SMEM_LIMIT = 232448 - 1024 - 256 - BIAS_SLOT_BYTES;  // 227 KB minus alignment slack, barriers, bias slots
constexpr int STAGING_BYTES = 32768;             // 8 epilogue warps x one 4 KB [32 rows][128 B] 128B-swizzled slot

struct GemmDev {
  int M, N, K;
  int splits, kb_per_split, num_kb;
  int m_blocks, n_blocks;
  void* out; int ld_out;
  void* out2; int ld_out2;
  const float* bias;
  const bf16* resid; int ld_resid;
  const bf16* aux; int ld_aux;
  float alpha;
  uint32_t flags;
  uint32_t drop_thresh16; float drop_scale; uint64_t seed; uint32_t site;
  unsigned long long* dbg;  // optional stall counters, 8 per CTA (merlot_gemm_debug_counters)
};

template <int BN, int EPI>
struct GemmCfg {
  static constexpr bool TS = EPI != 0;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (SMEM_LIMIT - (TS ? STAGING_BYTES : 0)) / STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;  // two accumulator stages, power-of-two allocation
  static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + (TS ? STAGING_BYTES : 0) + 1024 + 256 + BIAS_SLOT_BYTES;
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Fused epilogue math for 8 consecutive columns [col, col+8) of one row, in registers.
//   v = alpha*acc (+bias); GELU: pre <- v, v <- gelu(v); MUL_DGELU: v *= gelu'(aux); DROPOUT; (+resid)
// `in_range` = row < M (global loads are skipped for padding rows; their results are clipped on store).
__device__ __forceinline__ void epi_math8(const GemmDev& p, int row, int col, bool in_range, float (&v)[8], float (&pre)[8]) {
  const bool full = (col + 8 <= p.N);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] *= p.alpha;
  if (p.bias != nullptr) {
    if (full) {
      float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col));
      float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col + 4));
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < p.N) v[i] += __ldg(p.bias + col + i);
    }
  }
  if (p.flags & MERLOT_GEMM_GELU) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      pre[i] = (p.flags & MERLOT_GEMM_GELU_GRAD_OUT) ? gelu_erf_grad_fast(v[i]) : v[i];
      v[i] = gelu_erf_fast(v[i]);
    }
  }
  if ((p.flags & MERLOT_GEMM_MUL_AUX) && in_range) {
    const bf16* a = p.aux + (size_t)row * p.ld_aux + col;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (col + i < p.N) v[i] *= __bfloat162float(a[i]);
  }
  if ((p.flags & MERLOT_GEMM_MUL_DGELU) && in_range) {
    const bf16* a = p.aux + (size_t)row * p.ld_aux + col;
    if (full) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(a));
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      v[0] *= gelu_erf_grad_fast(f0.x); v[1] *= gelu_erf_grad_fast(f0.y); v[2] *= gelu_erf_grad_fast(f1.x);
      v[3] *= gelu_erf_grad_fast(f1.y); v[4] *= gelu_erf_grad_fast(f2.x); v[5] *= gelu_erf_grad_fast(f2.y);
      v[6] *= gelu_erf_grad_fast(f3.x); v[7] *= gelu_erf_grad_fast(f3.y);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < p.N) v[i] *= gelu_erf_grad_fast(__bfloat162float(a[i]));
    }
  }
  if (p.flags & MERLOT_GEMM_DROPOUT) {
    uint64_t lin = (uint64_t)row * (uint64_t)p.N + (uint64_t)col;  // col % 8 == 0, N % 8 == 0 enforced on host
    uint32_t keep = dropout_keep8(p.seed, p.site, lin >> 3, p.drop_thresh16);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = ((keep >> i) & 1u) ? v[i] * p.drop_scale : 0.0f;
  }
  if (p.resid != nullptr && in_range) {
    const bf16* r = p.resid + (size_t)row * p.ld_resid + col;
    if (full) {
      uint4 u = __ldg(reinterpret_cast<const uint4*>(r));
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      v[0] += f0.x; v[1] += f0.y; v[2] += f1.x; v[3] += f1.y; v[4] += f2.x; v[5] += f2.y; v[6] += f3.x; v[7] += f3.y;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < p.N) v[i] += __bfloat162float(r[i]);
    }
  }
}

__device__ __forceinline__ void store_bf16x8(bf16* o, int col, int N, const float (&v)[8]) {
  if (col + 8 <= N) {
    *reinterpret_cast<uint4*>(o) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (col + i < N) o[i] = __float2bfloat16_rn(v[i]);
  }
}

// direct (register -> global) store path: fp32 outputs, atomics, and bf16 fallbacks
__device__ __forceinline__ void epi_store_direct(const GemmDev& p, int row, int col, const float (&v)[8], const float (&pre)[8]) {
  const bool full = (col + 8 <= p.N);
  const bool dual = (p.flags & MERLOT_GEMM_GELU) && p.out2 != nullptr;
  if (p.flags & MERLOT_GEMM_OUT_F32) {
    float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ld_out + col;
    if (p.flags & MERLOT_GEMM_ATOMIC) {
      if (full) {
        red_add_v4(o, v[0], v[1], v[2], v[3]);
        red_add_v4(o + 4, v[4], v[5], v[6], v[7]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (col + i < p.N) atomicAdd(o + i, v[i]);
      }
    } else if (full && ((p.ld_out & 3) == 0)) {
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (col + i < p.N) o[i] = v[i];
    }
  } else if (dual) {
    store_bf16x8(reinterpret_cast<bf16*>(p.out) + (size_t)row * p.ld_out + col, col, p.N, pre);
    store_bf16x8(reinterpret_cast<bf16*>(p.out2) + (size_t)row * p.ld_out2 + col, col, p.N, v);
  } else {
    store_bf16x8(reinterpret_cast<bf16*>(p.out) + (size_t)row * p.ld_out + col, col, p.N, v);
  }
}



__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
  f[0] = f0.x; f[1] = f0.y; f[2] = f1.x; f[3] = f1.y; f[4] = f2.x; f[5] = f2.y; f[6] = f3.x; f[7] = f3.y;
}

// Epilogue feature masks.  The staged epilogue is instantiated once per feature set that the model uses (plus one generic,
// run-time-flag instance): with one generic body the plain path hops over the gelu / dropout / residual blocks and spends
// most of its time in instruction-cache misses (ncu: stall_no_inst on every BSSY/BSYNC of the skipped blocks).
enum : int { F_ALPHA = 1, F_BIAS = 2, F_GELU = 4, F_DUAL = 8, F_DGELU = 16, F_DROP = 32, F_RESID = 64, F_GRADOUT = 128, F_MULAUX = 256, F_GENERIC = -1 };
__device__ __forceinline__ int epi_features(const GemmDev& p) {
  return (p.alpha != 1.0f ? F_ALPHA : 0) | (p.bias ? F_BIAS : 0) | ((p.flags & MERLOT_GEMM_GELU) ? F_GELU : 0) |
         (((p.flags & MERLOT_GEMM_GELU) && p.out2) ? F_DUAL : 0) | ((p.flags & MERLOT_GEMM_MUL_DGELU) ? F_DGELU : 0) |
         ((p.flags & MERLOT_GEMM_DROPOUT) ? F_DROP : 0) | (p.resid ? F_RESID : 0) |
         ((p.flags & MERLOT_GEMM_GELU_GRAD_OUT) ? F_GRADOUT : 0) | ((p.flags & MERLOT_GEMM_MUL_AUX) ? F_MULAUX : 0);
}
template <int FL>
__device__ __forceinline__ bool feat(int run_time_features, int f) { return FL == F_GENERIC ? (run_time_features & f) != 0 : (FL & f) != 0; }

// fused math for 8 consecutive columns held in registers (chunk-local companion data already loaded)
template <int FL>
__device__ __forceinline__ void epi_math8_regs(const GemmDev& p, int rtf, int row, int col, bool in_range, float (&v)[8],
                                               uint32_t (&pre_packed)[4], const uint4& compq, const bf16* resid2,
                                               const float* bias8) {
  if (feat<FL>(rtf, F_ALPHA)) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= p.alpha;
  }
  if (feat<FL>(rtf, F_BIAS)) {  // this chunk's bias values sit in the warp's smem slot (zeros right of N): broadcast reads
    const float4 b0 = *reinterpret_cast<const float4*>(bias8);
    const float4 b1 = *reinterpret_cast<const float4*>(bias8 + 4);
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (feat<FL>(rtf, F_GELU)) {
    if (feat<FL>(rtf, F_GRADOUT)) {  // first output = gelu'(pre) (the FFN2 dgrad's factor), second = gelu(pre): Phi and exp shared
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float e0, e1;
        const float x0 = v[2 * k], x1 = v[2 * k + 1];
        const float c0 = normal_cdf_fast(x0, &e0), c1 = normal_cdf_fast(x1, &e1);
        pre_packed[k] = pack_bf16x2(fmaf(x0 * 0.39894228040143267794f, e0, c0), fmaf(x1 * 0.39894228040143267794f, e1, c1));
        v[2 * k] = x0 * c0;
        v[2 * k + 1] = x1 * c1;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) pre_packed[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = gelu_erf_fast(v[k]);
    }
  }
  if (feat<FL>(rtf, F_DGELU)) {  // companion = gelu' input
    float a[8];
    unpack8(compq, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= gelu_erf_grad_fast(a[k]);
  }
  if (feat<FL>(rtf, F_MULAUX)) {  // companion = the saved factor
    float a[8];
    unpack8(compq, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= a[k];
  }
  if (feat<FL>(rtf, F_DROP)) {
    const uint64_t lin = (uint64_t)row * (uint64_t)p.N + (uint64_t)col;  // col % 8 == 0, N % 8 == 0 enforced on host
    const uint32_t keep = dropout_keep8(p.seed, p.site, lin >> 3, p.drop_thresh16);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = ((keep >> k) & 1u) ? v[k] * p.drop_scale : 0.0f;
  }
  if (feat<FL>(rtf, F_RESID)) {
    if (resid2 == nullptr) {  // companion = residual
      float a[8];
      unpack8(compq, a);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] += a[k];
    } else if (in_range && col < p.N) {  // residual next to a dual output or a gelu' companion: rare, unprefetched
      if (col + 8 <= p.N) {
        float a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(resid2 + col)), a);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += a[k];
      } else {
#pragma unroll 1
        for (int k = 0; k < 8; ++k)
          if (col + k < p.N) v[k] += __bfloat162float(resid2[col + k]);
      }
    }
  }
}

// Coalesced companion fetch: this warp's [32 rows][64 cols] bf16 block, lane l holding 16 B of row 4k + l/8, chunk l%8.
__device__ __forceinline__ void comp_fetch(const bf16* base, int ld, int row0q, int ucol0, int M, int N, int lane, uint4 (&cp)[8]) {
  const int cc = lane & 7;
  const int gcol = ucol0 + cc * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int grow = row0q + k * 4 + (lane >> 3);
    if (grow < M && gcol + 8 <= N) {
      cp[k] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)grow * ld + gcol));
    } else {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (grow < M)
        for (int i = 0; i < 8; ++i)
          if (gcol + i < N) w[i >> 1] |= (uint32_t)__bfloat16_as_ushort(base[(size_t)grow * ld + gcol + i]) << ((i & 1) * 16);
      cp[k] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}
static __device__ __noinline__ void store_bf16_tail(bf16* o, uint4 q, int n) {  // n < 8 trailing columns (cold path)
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  for (int i = 0; i < n; ++i) o[i] = __ushort_as_bfloat16((unsigned short)(w[i >> 1] >> ((i & 1) * 16)));
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp-private staged epilogue (EPI 1 and 2), shared by the 1-CTA and the CTA-pair kernel.
// Every epilogue warp owns the 32 accumulator rows of its TMEM lane quadrant, every second column unit of the tile and a
// private 4 KB smem slot ([32 rows][128 B], 128B-swizzled).  No CTA-wide barrier and no TMA on the way out (a bulk store
// queues behind the main loop's loads in the SM's TMA FIFO and comes back thousands of cycles later -- measured):
//   * the accumulator chunk is loaded one tcgen05.ld ahead of the math;
//   * the thread that owns row `lane` writes its packed results into the slot; the warp then drains the slot with
//     row-contiguous 16-byte accesses (4 rows x 128 B per instruction): full-line st.global / red.global.add.v4.f32;
//   * the companion operand (gelu' input or residual) takes the same road backwards: coalesced loads one unit ahead,
//     parked in the slot, read back by the owning thread and overwritten in place by the result.
// `release()` is called by the whole warp once the last tcgen05.ld of the tile has completed (TMEM stage reusable).
// ---------------------------------------------------------------------------------------------------------------------
// Operands fetched ahead of their use and carried from one tile's epilogue into the next (registers; the tile loop is inlined)
struct EpiCarry {
  uint4 cp[8];  // companion block of the next 64-column unit
  float bnext;  // bias value of column (next chunk) + lane
};

template <int BN, int EPI, int FL>
struct EpiPlan {  // which chunks / columns an epilogue warp walks (depends only on the feature set and the warp's half)
  bool dual, wide;
  int half, my_chunks;
  __device__ __forceinline__ EpiPlan(int rtf, int half_) : half(half_) {
    dual = EPI == 1 && feat<FL>(rtf, F_DUAL);
    wide = EPI == 1 && !dual;
    constexpr int NU0 = (BN / 64 + 1) / 2;  // 64-column units of half 0 (half 1 owns BN/64 - NU0)
    my_chunks = wide ? 2 * (half == 0 ? NU0 : BN / 64 - NU0) : BN / 64;
  }
  __device__ __forceinline__ int tcol_of(int i) const { return wide ? ((2 * (i >> 1) + half) * 2 + (i & 1)) * 32 : (2 * i + half) * 32; }
};

template <int BN, int EPI, int FL>
__device__ __forceinline__ void epilogue_prefetch(const GemmDev& p, int rtf, int row0q, int n0, int half, int lane, EpiCarry& cy) {
  const EpiPlan<BN, EPI, FL> plan(rtf, half);
  if (EPI == 1 && feat<FL>(rtf, F_BIAS)) {
    const int c = n0 + plan.tcol_of(0) + lane;
    cy.bnext = c < p.N ? __ldg(p.bias + c) : 0.0f;
  }
  if (plan.wide && (feat<FL>(rtf, F_DGELU) || feat<FL>(rtf, F_MULAUX) || feat<FL>(rtf, F_RESID))) {
    const bool dg = feat<FL>(rtf, F_DGELU) || feat<FL>(rtf, F_MULAUX);
    comp_fetch(dg ? p.aux : p.resid, dg ? p.ld_aux : p.ld_resid, row0q, n0 + plan.tcol_of(0), p.M, p.N, lane, cy.cp);
  }
}

// One tile.  `cy` holds the operands prefetched for this tile's first chunk (epilogue_prefetch, issued by the caller before it
// waits for the accumulator so that the wait hides their latency).
template <int BN, int EPI, int FL, typename Release>
__device__ __forceinline__ void epilogue_tile_loop(const GemmDev& p, int rtf, uint8_t* slot, float* bias_slot, uint32_t taddr, int row0q, int n0,
                                                   int half, int lane, EpiCarry& cy, Release&& release) {
  // the slots are shared memory; without the hint the pointers (through the lambda/struct plumbing) compile to generic LD.E/ST.E
  __builtin_assume(__isShared(slot));
  __builtin_assume(__isShared(bias_slot));
  // EPI 1, single output: 64-column units of two chunks (unit index 2*j + half).  Dual output and fp32: one 32-column chunk per
  // unit (index 2*i + half); the dual slot row is [pre 64 B | act 64 B].
  // The chunk loop is deliberately NOT unrolled (one copy of the math body): a fully unrolled epilogue is 20k instructions and
  // thrashes the instruction cache (measured 2.5x slower).  The accumulator double buffer is a register copy instead.
  const EpiPlan<BN, EPI, FL> plan(rtf, half);
  const bool dual = plan.dual, wide = plan.wide;
  const int my_chunks = plan.my_chunks;
  const int row = row0q + lane;
  const bool in_range = row < p.M;
  const bool do_dgelu = feat<FL>(rtf, F_DGELU) || feat<FL>(rtf, F_MULAUX);  // the companion operand is p.aux
  const bf16* comp = !wide ? nullptr : do_dgelu ? p.aux : (feat<FL>(rtf, F_RESID) ? p.resid : nullptr);
  const int ld_comp = do_dgelu ? p.ld_aux : p.ld_resid;
  const bf16* resid2 = ((!wide || do_dgelu) && feat<FL>(rtf, F_RESID)) ? p.resid + (size_t)row * p.ld_resid : nullptr;  // rare: unprefetched
  const bool has_bias = EPI == 1 && feat<FL>(rtf, F_BIAS);
  uint32_t rn[32], rc[32];
  tmem_ld_32x32(taddr + (uint32_t)plan.tcol_of(0), rn);
#pragma unroll 1
  for (int i = 0; i < my_chunks; ++i) {
    const int tcol = plan.tcol_of(i);
    const int col0 = n0 + tcol;
    const int sub = wide ? (i & 1) : 0;
    const int ucol0 = wide ? n0 + plan.tcol_of(i & ~1) : col0;  // first column of the unit
    const bool live = ucol0 < p.N;                              // unit not entirely right of the matrix (warp-uniform)
    if (comp && sub == 0) {
      // park the companion block in the slot, then start fetching the next unit's (or the next tile's first)
      if (live) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<uint4*>(slot + sw128_offset((uint32_t)(k * 4 + (lane >> 3)), (uint32_t)(lane & 7))) = cy.cp[k];
      }
      if (i + 2 < my_chunks) comp_fetch(comp, ld_comp, row0q, n0 + plan.tcol_of(i + 2), p.M, p.N, lane, cy.cp);
      __syncwarp();
    }
    if (has_bias) {
      __syncwarp();  // the chunk before has finished reading the slot
      bias_slot[lane] = cy.bnext;
      __syncwarp();
      const int c = n0 + plan.tcol_of(i + 1) + lane;
      cy.bnext = (i + 1 < my_chunks && c < p.N) ? __ldg(p.bias + c) : 0.0f;
    }
    tmem_wait_ld_regs(rn);
#pragma unroll
    for (int k = 0; k < 32; ++k) rc[k] = rn[k];
    if (i + 1 < my_chunks) {
      tmem_ld_32x32(taddr + (uint32_t)plan.tcol_of(i + 1), rn);
    } else {
      release();
    }
    if (!live) continue;
    if (EPI == 2) {
#pragma unroll
      for (int g = 0; g < 8; ++g)
        *reinterpret_cast<uint4*>(slot + sw128_offset((uint32_t)lane, (uint32_t)g)) =
            make_uint4(__float_as_uint(__uint_as_float(rc[g * 4 + 0]) * p.alpha), __float_as_uint(__uint_as_float(rc[g * 4 + 1]) * p.alpha),
                       __float_as_uint(__uint_as_float(rc[g * 4 + 2]) * p.alpha), __float_as_uint(__uint_as_float(rc[g * 4 + 3]) * p.alpha));
      __syncwarp();
      const int cc = lane & 7;
      const int gcol = col0 + cc * 4;
      float* obase = reinterpret_cast<float*>(p.out);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = k * 4 + (lane >> 3);
        const int grow = row0q + rr;
        const float4 q = *reinterpret_cast<const float4*>(slot + sw128_offset((uint32_t)rr, (uint32_t)cc));
        if (grow < p.M && gcol < p.N) {
          float* o = obase + (size_t)grow * p.ld_out + gcol;
          if (gcol + 4 <= p.N) {
            red_add_v4(o, q.x, q.y, q.z, q.w);
          } else {
            if (gcol + 0 < p.N) atomicAdd(o + 0, q.x);
            if (gcol + 1 < p.N) atomicAdd(o + 1, q.y);
            if (gcol + 2 < p.N) atomicAdd(o + 2, q.z);
          }
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int col = col0 + g * 8;
        float v[8];
        uint32_t pre_packed[4];
        uint4 compq = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t off = sw128_offset((uint32_t)lane, (uint32_t)(sub * 4 + g));
        if (comp) compq = *reinterpret_cast<const uint4*>(slot + off);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(rc[g * 8 + k]);
        epi_math8_regs<FL>(p, rtf, row, col, in_range, v, pre_packed, compq, resid2, bias_slot + g * 8);
        const uint4 outq = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        if (dual) {
          *reinterpret_cast<uint4*>(slot + off) = make_uint4(pre_packed[0], pre_packed[1], pre_packed[2], pre_packed[3]);
          *reinterpret_cast<uint4*>(slot + sw128_offset((uint32_t)lane, (uint32_t)(4 + g))) = outq;
        } else {
          *reinterpret_cast<uint4*>(slot + off) = outq;
        }
      }
      if (dual || sub == 1) {
        __syncwarp();
        // dual: chunks 0-3 of every slot row -> out (pre-activation), chunks 4-7 -> out2 (activation); single: all 8 -> out
        const int cc = lane & 7;
        const bool second = dual && cc >= 4;
        bf16* obase = reinterpret_cast<bf16*>(second ? p.out2 : p.out);
        const int ldo = second ? p.ld_out2 : p.ld_out;
        const int gcol = ucol0 + (dual ? (cc & 3) : cc) * 8;
        if (gcol < p.N) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = k * 4 + (lane >> 3);
            const int grow = row0q + rr;
            if (grow < p.M) {
              const uint4 q = *reinterpret_cast<const uint4*>(slot + sw128_offset((uint32_t)rr, (uint32_t)cc));
              bf16* o = obase + (size_t)grow * ldo + gcol;
              if (gcol + 8 <= p.N) *reinterpret_cast<uint4*>(o) = q;
              else store_bf16_tail(o, q, p.N - gcol);
            }
          }
        }
        __syncwarp();
      }
    }
  }
}

// bench.py roofline hook (gemm_tcgen05.cu): CUDA events around a K1 launch when profiling is switched on
void* gemm_prof_before(double flops, cudaStream_t stream);
void gemm_prof_after(void* tok, cudaStream_t stream);

}  // namespace mb
