// K6/K7 and glue: token-assembly kernels around the three transformer stacks (single HBM pass each, 128-bit accesses).
//   * patch im2col                 utils/vision_transformer.py:193-205 (x - 0.5, 16x16/16 VALID conv as a GEMM operand)
//   * ViT token assembly           utils/vision_transformer.py:229-233 (+2 zero CLS slots, + position_embedder2d)
//   * viz assembly ("K7")          utils/vision_transformer.py:251-266 (cls/seq split, 2x2 avg-pool) +
//                                  model/modeling.py:99-125,299-337 (cls0 || pooled seq, + img_idx_pe + final_pe)
//   * word embedding ("K6")        model/modeling.py:262-292 (E[ids] + Pos[0:L]) via utils/model_utils.py:238-310
//   * grouped row sums             gradients of the broadcast position tables
// The LayerNorms that follow each assembly are the generic K5 kernels (rowwise.cu).
#include "host_common.h"
#include "ptx.cuh"

namespace mb {

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// image bf16 NHWC [N,H0,W0,3] -> A bf16 [N*h1*w1, P*P*3], A = bf16(x - 0.5); one thread per 8 output elements
__global__ void im2col_kernel(const bf16* __restrict__ img, bf16* __restrict__ a, int N, int H0, int W0, int P) {
  const int h1 = H0 / P, w1 = W0 / P, rowlen = P * 3, K = P * rowlen, per_kh = rowlen / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * h1 * w1 * P * per_kh;
  if (gid >= total) return;
  const int c8 = (int)(gid % per_kh);
  long long t = gid / per_kh;
  const int kh = (int)(t % P); t /= P;
  const int j = (int)(t % w1); t /= w1;
  const int i = (int)(t % h1);
  const int n = (int)(t / h1);
  const size_t src = (((size_t)n * H0 + (size_t)i * P + kh) * W0 + (size_t)j * P) * 3 + (size_t)c8 * 8;
  float v[8];
  ld8(img + src, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] -= 0.5f;
  const size_t row = ((size_t)n * h1 + i) * w1 + j;
  st8(a + row * K + (size_t)kh * rowlen + (size_t)c8 * 8, v);
}

// xsum[n, t] = (t < ncls ? 0 : patch[n, t-ncls]) + pe(t);  pos table row of patch (i,j) is i*tab_w + j
__global__ void vit_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ pos, const float* __restrict__ cls,
                                    float* __restrict__ xsum, int N, int h1, int w1, int ncls, int tab_w, int H) {
  const int per = H / 8, Sv = h1 * w1 + ncls;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)N * Sv * per) return;
  const int c = (int)(gid % per) * 8;
  const long long r = gid / per;
  const int t = (int)(r % Sv), n = (int)(r / Sv);
  float v[8], e[8];
  if (t < ncls) {
    ld8(cls + (size_t)t * H + c, v);
  } else {
    const int pidx = t - ncls, i = pidx / w1, j = pidx % w1;
    ld8(patch + ((size_t)n * h1 * w1 + pidx) * H + c, v);
    ld8(pos + ((size_t)i * tab_w + j) * H + c, e);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += e[k];
  }
  st8(xsum + (size_t)r * H + c, v);
}
// d_patch (bf16, wgrad operand) = dxsum[:, ncls:]
__global__ void vit_assemble_bwd_kernel(const float* __restrict__ dxsum, bf16* __restrict__ dpatch, int N, int np, int ncls, int H) {
  const int per = H / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)N * np * per) return;
  const int c = (int)(gid % per) * 8;
  const long long r = gid / per;
  const int pidx = (int)(r % np), n = (int)(r / np);
  float v[8];
  ld8(dxsum + ((size_t)n * (np + ncls) + ncls + pidx) * H + c, v);
  st8(dpatch + (size_t)r * H + c, v);
}

// K7 forward.  hv bf16 [N*Sv, H] (ViT output after ln_final); xsum fp32 [B*P, H], P = ncg*vcl, vcl = h2*w2+1;
// img_trg fp32 [N, H] = hv[:, 1] (contrastive target, modeling.py:99)
__global__ void viz_assemble_kernel(const bf16* __restrict__ hv, const float* __restrict__ img_idx_pe, const int* __restrict__ img_idx,
                                    const float* __restrict__ fpos, const float* __restrict__ fcls, float* __restrict__ xsum,
                                    float* __restrict__ img_trg, int N, int h1, int w1, int ncls, int sp, int tab_w, int H) {
  const int per = H / 8, h2 = h1 / sp, w2 = w1 / sp, vcl = h2 * w2 + 1, Sv = h1 * w1 + ncls;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)N * (vcl + 1) * per) return;
  const int c = (int)(gid % per) * 8;
  const long long r = gid / per;
  const int t = (int)(r % (vcl + 1)), n = (int)(r / (vcl + 1));
  const bf16* base = hv + (size_t)n * Sv * H + c;
  float v[8];
  if (t == vcl) {  // extra slot: contrastive target = second CLS token
    ld8(base + (size_t)1 * H, v);
    st8(img_trg + (size_t)n * H + c, v);
    return;
  }
  float e[8];
  if (t == 0) {
    ld8(base, v);
    ld8(fcls + c, e);
  } else {
    const int i2 = (t - 1) / w2, j2 = (t - 1) % w2;
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
    for (int a = 0; a < sp; ++a)
      for (int b = 0; b < sp; ++b) {
        float u[8];
        ld8(base + (size_t)(ncls + (i2 * sp + a) * w1 + (j2 * sp + b)) * H, u);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] += u[k];
      }
    const float inv = 1.f / (float)(sp * sp);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= inv;
    ld8(fpos + ((size_t)i2 * tab_w + j2) * H + c, e);
  }
  float g[8];
  ld8(img_idx_pe + (size_t)img_idx[n] * H + c, g);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] += e[k] + g[k];
  st8(xsum + ((size_t)n * vcl + t) * H + c, v);  // frames are consecutive inside a group: row = b*P + s*vcl + t = n*vcl + t
}
// K7 backward: d_hv bf16 [N*Sv, H] from dxsum fp32 [N*vcl, H] and d_img_trg fp32 [N, H]
__global__ void viz_assemble_bwd_kernel(const float* __restrict__ dxsum, const float* __restrict__ d_img_trg, bf16* __restrict__ dhv,
                                        int N, int h1, int w1, int ncls, int sp, int H) {
  const int per = H / 8, h2 = h1 / sp, w2 = w1 / sp, vcl = h2 * w2 + 1, Sv = h1 * w1 + ncls;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)N * Sv * per) return;
  const int c = (int)(gid % per) * 8;
  const long long r = gid / per;
  const int t = (int)(r % Sv), n = (int)(r / Sv);
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] = 0.f;
  if (t == 0) {
    ld8(dxsum + ((size_t)n * vcl) * H + c, v);
  } else if (t == 1 && ncls > 1) {
    if (d_img_trg) ld8(d_img_trg + (size_t)n * H + c, v);
  } else if (t >= ncls) {
    const int pidx = t - ncls, i = pidx / w1, j = pidx % w1;
    if (i < h2 * sp && j < w2 * sp) {
      ld8(dxsum + ((size_t)n * vcl + 1 + (i / sp) * w2 + (j / sp)) * H + c, v);
      const float inv = 1.f / (float)(sp * sp);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= inv;
    }
  }
  st8(dhv + (size_t)r * H + c, v);
}

// K6 forward: xsum[r] = E[ids[r]] + Pos[r % L]
__global__ void embed_kernel(const int* __restrict__ ids, const float* __restrict__ emb, const float* __restrict__ pos,
                             float* __restrict__ xsum, long long R, int L, int H) {
  const int per = H / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= R * per) return;
  const int c = (int)(gid % per) * 8;
  const long long r = gid / per;
  float v[8], e[8];
  ld8(emb + (size_t)ids[r] * H + c, v);
  ld8(pos + (size_t)(r % L) * H + c, e);
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] += e[k];
  st8(xsum + (size_t)r * H + c, v);
}

// dst[idxmap[t]] += sum_g src[g*per + t0 + t], t in [0, nt)  -- gradient of a table broadcast over `groups` sequences
__global__ void group_rowsum_kernel(const float* __restrict__ src, int ld, int groups, int per, int t0, int nt,
                                    const int* __restrict__ idxmap, float* __restrict__ dst, int ld_dst, int H) {
  const int perc = H / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)nt * perc) return;
  const int c = (int)(gid % perc) * 8;
  const int t = (int)(gid / perc);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int g = 0; g < groups; ++g) {
    float v[8];
    ld8(src + ((size_t)g * per + t0 + t) * ld + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
  }
  float* d = dst + (size_t)(idxmap ? idxmap[t] : t) * ld_dst + c;
  float o[8];
  ld8(d, o);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] += acc[k];
  st8(d, o);
}
// dst[idx[n]] += sum_{t<per} src[n*per + t]   (img_idx_pe gradient; idx may repeat -> atomics)
__global__ void segment_rowsum_scatter_kernel(const float* __restrict__ src, int ld, int n_seg, int per, const int* __restrict__ idx,
                                              float* __restrict__ dst, int ld_dst, int H) {
  const int perc = H / 8;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n_seg * perc) return;
  const int c = (int)(gid % perc) * 8;
  const int n = (int)(gid / perc);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = 0; t < per; ++t) {
    float v[8];
    ld8(src + ((size_t)n * per + t) * ld + c, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += v[k];
  }
  float* d = dst + (size_t)idx[n] * ld_dst + c;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3]) : "memory");
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(acc[4]), "f"(acc[5]), "f"(acc[6]), "f"(acc[7]) : "memory");
}

}  // namespace mb

using namespace mb;

#define GRID1D(n) (unsigned)ceil_div_ll((n), 256), 256, 0, st

extern "C" int merlot_patch_im2col(const void* image, void* a, int N, int H0, int W0, int P, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(image && a, MERLOT_EINVAL, "patch_im2col: null pointer");
  MB_REQUIRE(P > 0 && H0 % P == 0 && W0 % P == 0, MERLOT_ESHAPE, "patch_im2col: image %dx%d not divisible by patch %d", H0, W0, P);
  MB_REQUIRE((P * 3) % 8 == 0, MERLOT_ESHAPE, "patch_im2col: patch_size*3 must be a multiple of 8");
  const long long total = (long long)N * (H0 / P) * (W0 / P) * P * (P * 3 / 8);
  im2col_kernel<<<GRID1D(total)>>>((const bf16*)image, (bf16*)a, N, H0, W0, P);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

extern "C" int merlot_vit_assemble_fwd(const float* patch, const float* pos_table, const float* cls_emb, float* xsum, int N,
                                       int h1, int w1, int ncls, int tab_w, int H, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(patch && pos_table && xsum && (ncls == 0 || cls_emb), MERLOT_EINVAL, "vit_assemble_fwd: null pointer");
  MB_REQUIRE(H % 8 == 0 && h1 <= tab_w && w1 <= tab_w, MERLOT_ESHAPE, "vit_assemble_fwd: H %% 8 != 0 or grid %dx%d exceeds table %d", h1, w1, tab_w);
  vit_assemble_kernel<<<GRID1D((long long)N * (h1 * w1 + ncls) * (H / 8))>>>(patch, pos_table, cls_emb, xsum, N, h1, w1, ncls, tab_w, H);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}
extern "C" int merlot_vit_assemble_bwd(const float* dxsum, void* dpatch, int N, int np, int ncls, int H, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(dxsum && dpatch, MERLOT_EINVAL, "vit_assemble_bwd: null pointer");
  vit_assemble_bwd_kernel<<<GRID1D((long long)N * np * (H / 8))>>>(dxsum, (bf16*)dpatch, N, np, ncls, H);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

extern "C" int merlot_viz_assemble_fwd(const void* hv, const float* img_idx_pe, const int* img_idx, const float* final_pos,
                                       const float* final_cls, float* xsum, float* img_trg, int N, int h1, int w1, int ncls,
                                       int sp, int tab_w, int H, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(hv && img_idx_pe && img_idx && final_pos && final_cls && xsum && img_trg, MERLOT_EINVAL, "viz_assemble_fwd: null pointer");
  MB_REQUIRE(H % 8 == 0 && sp >= 1 && ncls >= 2, MERLOT_ESHAPE, "viz_assemble_fwd: need H %% 8 == 0, spatial_pool_size >= 1, num_cls_emb >= 2");
  const int vcl = (h1 / sp) * (w1 / sp) + 1;
  viz_assemble_kernel<<<GRID1D((long long)N * (vcl + 1) * (H / 8))>>>((const bf16*)hv, img_idx_pe, img_idx, final_pos, final_cls, xsum,
                                                                       img_trg, N, h1, w1, ncls, sp, tab_w, H);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}
extern "C" int merlot_viz_assemble_bwd(const float* dxsum, const float* d_img_trg, void* dhv, int N, int h1, int w1, int ncls,
                                       int sp, int H, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(dxsum && dhv, MERLOT_EINVAL, "viz_assemble_bwd: null pointer");
  viz_assemble_bwd_kernel<<<GRID1D((long long)N * (h1 * w1 + ncls) * (H / 8))>>>(dxsum, d_img_trg, (bf16*)dhv, N, h1, w1, ncls, sp, H);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

extern "C" int merlot_embed_fwd(const int* ids, const float* emb, const float* pos, float* xsum, long long R, int L, int H,
                                void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(ids && emb && pos && xsum, MERLOT_EINVAL, "embed_fwd: null pointer");
  MB_REQUIRE(H % 8 == 0 && L > 0, MERLOT_ESHAPE, "embed_fwd: H %% 8 != 0 or L <= 0");
  if (R == 0) return MERLOT_OK;
  embed_kernel<<<GRID1D(R * (H / 8))>>>(ids, emb, pos, xsum, R, L, H);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}

extern "C" int merlot_group_rowsum(const float* src, int ld, int groups, int per, int t0, int nt, const int* idxmap, float* dst,
                                   int ld_dst, int H, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(src && dst, MERLOT_EINVAL, "group_rowsum: null pointer");
  MB_REQUIRE(H % 8 == 0 && t0 >= 0 && t0 + nt <= per, MERLOT_ESHAPE, "group_rowsum: bad range");
  if (nt == 0 || groups == 0) return MERLOT_OK;
  group_rowsum_kernel<<<GRID1D((long long)nt * (H / 8))>>>(src, ld, groups, per, t0, nt, idxmap, dst, ld_dst, H);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}
extern "C" int merlot_segment_rowsum_scatter(const float* src, int ld, int n_seg, int per, const int* idx, float* dst, int ld_dst,
                                             int H, void* stream_) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_);
  MB_REQUIRE(src && dst && idx, MERLOT_EINVAL, "segment_rowsum_scatter: null pointer");
  MB_REQUIRE(H % 8 == 0, MERLOT_ESHAPE, "segment_rowsum_scatter: H %% 8 != 0");
  if (n_seg == 0) return MERLOT_OK;
  segment_rowsum_scatter_kernel<<<GRID1D((long long)n_seg * (H / 8))>>>(src, ld, n_seg, per, idx, dst, ld_dst, H);
  MB_CHECK_LAUNCH();
  return MERLOT_OK;
}
