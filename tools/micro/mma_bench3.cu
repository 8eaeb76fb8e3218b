// Microbenchmark 3: tcgen05.mma issue floor with compile-time shapes, descriptors advanced by a constant, warp-uniform issue.
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

template <int M, int N, bool TS, int REPS, bool DEP, int AMN = 0, int BMN = 0>
__device__ __forceinline__ void run_case(uint32_t tmem, uint32_t aa, uint32_t ba, uint64_t* bar, uint32_t& phase, bool leader, long long* out) {
  constexpr uint32_t idesc = make_idesc_bf16(M, N, AMN, BMN);
  const uint64_t da0 = AMN ? desc_mnmajor(aa, 0, 0) : desc_kmajor(aa, 0), db0 = BMN ? desc_mnmajor(ba, 0, 0) : desc_kmajor(ba, 0);
  for (int warm = 0; warm < 2; ++warm) {
    const long long t0 = clock64();
    if (leader) {
#pragma unroll
      for (int r = 0; r < REPS; ++r) {
        const uint64_t da = da0 + (uint64_t)((r & 3) * (AMN ? 128 : 2)), db = db0 + (uint64_t)((r & 3) * (BMN ? 128 : 2));
        const uint32_t d = tmem + (DEP ? 0u : (uint32_t)((r & 1) * 256));
        if (TS) umma_bf16_ts(d, tmem + 384, db, idesc, 1u);
        else umma_bf16_ss(d, da, db, idesc, 1u);
      }
    }
    const long long t1 = clock64();
    if (leader) umma_commit(bar);
    mbar_wait(bar, phase);
    phase ^= 1;
    const long long t2 = clock64();
    if (warm == 1 && leader) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = REPS; out[3] = M * 1000 + N + (TS ? 500000 : 0) + (DEP ? 0 : 1000000) + AMN * 10000000 + BMN * 20000000; }
  }
}

__global__ void __launch_bounds__(128, 1) bench(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_u == 0) {
    const bool leader = elect_one();
    uint32_t phase = 0;
    const uint32_t aa = smem_u32(smem), ba = smem_u32(smem + 16384);
    run_case<128, 256, false, 32, true>(tmem, aa, ba, bar, phase, leader, out + 0);
    run_case<128, 128, false, 32, true>(tmem, aa, ba, bar, phase, leader, out + 4);
    run_case<128, 64, false, 32, true>(tmem, aa, ba, bar, phase, leader, out + 8);
    run_case<128, 32, false, 32, true>(tmem, aa, ba, bar, phase, leader, out + 12);
    run_case<128, 16, false, 32, true>(tmem, aa, ba, bar, phase, leader, out + 16);
    run_case<128, 64, false, 32, false>(tmem, aa, ba, bar, phase, leader, out + 20);
    run_case<64, 64, false, 32, true>(tmem, aa, ba, bar, phase, leader, out + 24);
    run_case<128, 64, true, 32, true>(tmem, aa, ba, bar, phase, leader, out + 28);
    run_case<128, 128, true, 32, true>(tmem, aa, ba, bar, phase, leader, out + 32);
    run_case<128, 64, false, 4, true>(tmem, aa, ba, bar, phase, leader, out + 36);
    run_case<128, 64, false, 8, true>(tmem, aa, ba, bar, phase, leader, out + 40);
    run_case<128, 64, false, 1, true>(tmem, aa, ba, bar, phase, leader, out + 44);
    run_case<128, 192, false, 32, true>(tmem, aa, ba, bar, phase, leader, out + 48);
    run_case<128, 64, false, 32, true, 0, 1>(tmem, aa, ba, bar, phase, leader, out + 52);
    run_case<64, 64, false, 32, true, 1, 1>(tmem, aa, ba, bar, phase, leader, out + 56);
    run_case<128, 64, false, 32, true, 1, 1>(tmem, aa, ba, bar, phase, leader, out + 60);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* out;
  cudaMalloc(&out, 80 * 8);
  cudaMemset(out, 0, 80 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 52000);
  bench<<<1, 128, 52000>>>(out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  long long r[80];
  cudaMemcpy(r, out, 80 * 8, cudaMemcpyDeviceToHost);
  printf("code(M*1000+N, +500000 TS, +1000000 independent D, +10000000 A MN-major, +20000000 B MN-major)  reps  issue cyc/MMA   total/MMA   total\n");
  for (int c = 0; c < 16; ++c)
    printf("%8lld %5lld %10.1f %10.1f %8lld\n", r[c * 4 + 3], r[c * 4 + 2], (double)r[c * 4] / r[c * 4 + 2], (double)r[c * 4 + 1] / r[c * 4 + 2], r[c * 4 + 1]);
  return 0;
}
