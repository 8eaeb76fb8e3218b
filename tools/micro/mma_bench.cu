// Microbenchmark: cost of tcgen05.mma (kind::f16, bf16) per instruction on B200 as a function of the tile shape, operand
// majorness, accumulator dependence and A-operand source (smem descriptor vs TMEM).  One CTA, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I merlot_b200/csrc -I include tools/micro/mma_bench.cu -o tools/micro/mma_bench.bin
#include "ptx.cuh"
using namespace mb;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct Case { int M, N, a_mn, b_mn, dep, a_tmem, reps; };

__global__ void __launch_bounds__(128, 1) bench(const Case* cases, int ncases, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 16 KB
  uint8_t* sB = smem + 16384;    // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // small bf16 values
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    for (int c = 0; c < ncases; ++c) {
      const Case cs = cases[c];
      const uint32_t idesc = make_idesc_bf16(cs.M, cs.N, cs.a_mn, cs.b_mn);
      const uint32_t aa = smem_u32(sA), ba = smem_u32(sB);
      for (int warm = 0; warm < 2; ++warm) {
        const long long t0 = clock64();
        for (int r = 0; r < cs.reps; ++r) {
          const int k = r & 3;
          const uint64_t da = cs.a_mn ? desc_mnmajor(aa, k, 0) : desc_kmajor(aa, k);
          const uint64_t db = cs.b_mn ? desc_mnmajor(ba, k, 8192) : desc_kmajor(ba, k);
          const uint32_t d = tmem + (cs.dep ? 0u : (uint32_t)((r & 1) * 256));
          if (cs.a_tmem) umma_bf16_ts(d, tmem + 384, db, idesc, 1u);
          else umma_bf16_ss(d, da, db, idesc, 1u);
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, phase);
        phase ^= 1;
        const long long t2 = clock64();
        if (warm == 1) { out[c * 2] = t1 - t0; out[c * 2 + 1] = t2 - t0; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  Case h[] = {
      {128, 256, 0, 0, 1, 0, 64}, {128, 128, 0, 0, 1, 0, 64}, {128, 64, 0, 0, 1, 0, 64}, {128, 32, 0, 0, 1, 0, 64}, {128, 16, 0, 0, 1, 0, 64},
      {128, 64, 0, 0, 0, 0, 64}, {128, 128, 0, 0, 0, 0, 64},
      {128, 64, 0, 1, 1, 0, 64}, {128, 64, 1, 1, 1, 0, 64}, {64, 64, 1, 1, 1, 0, 64}, {64, 64, 0, 0, 1, 0, 64},
      {128, 64, 0, 0, 1, 1, 64}, {128, 64, 0, 1, 1, 1, 64}, {128, 128, 0, 0, 1, 1, 64},
      {128, 64, 0, 0, 1, 0, 1}, {128, 64, 0, 0, 1, 0, 4}, {128, 64, 0, 0, 1, 0, 8}, {128, 64, 0, 0, 1, 0, 16}, {128, 128, 0, 0, 1, 0, 4},
  };
  const int n = sizeof(h) / sizeof(h[0]);
  Case* d; long long* out;
  cudaMalloc(&d, sizeof(h)); cudaMalloc(&out, n * 16);
  cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  bench<<<1, 128, 52000>>>(d, n, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long r[64];
  cudaMemcpy(r, out, n * 16, cudaMemcpyDeviceToHost);
  printf("%-6s %-5s %-5s %-5s %-4s %-6s %-5s | issue cyc/MMA | total cyc (issue..mbarrier wake) | cyc/MMA\n", "M", "N", "A_mn", "B_mn", "dep", "A_tmem", "reps");
  for (int c = 0; c < n; ++c)
    printf("%-6d %-5d %-5d %-5d %-4d %-6d %-5d | %13.1f | %32lld | %7.1f\n", h[c].M, h[c].N, h[c].a_mn, h[c].b_mn, h[c].dep, h[c].a_tmem, h[c].reps,
           (double)r[c * 2] / h[c].reps, r[c * 2 + 1], (double)r[c * 2 + 1] / h[c].reps);
  return 0;
}
