// K1 1-CTA kernel instances for the 192-column tile (see gemm_kernel.cuh)
#include "gemm_kernel.cuh"

namespace mb {
template int launch_gemm_bn<192>(bool, bool, int, int, const CUtensorMap&, const CUtensorMap&, const GemmDev&, int, cudaStream_t);
}
