// ib_t32_phase.cu -- instantiations of the fused per-phase kernels of the |T| <= 32 family (ib_kernels_t32.cuh)
#include "ib_kernels_t32.cuh"
namespace ibldpc {
T32PhaseKernel t32_phase_cn_kernel(bool early) { return early ? ib_t32_phase_kernel<kPhaseCn, true> : ib_t32_phase_kernel<kPhaseCn, false>; }
T32PhaseKernel t32_phase_vn_kernel() { return ib_t32_phase_kernel<kPhaseVn, false>; }
T32PhaseKernel t32_phase_out_kernel() { return ib_t32_phase_kernel<kPhaseOut, false>; }
T32CoopKernel t32_coop_kernel(bool early) { return early ? ib_t32_coop_kernel<true> : ib_t32_coop_kernel<false>; }
}  // namespace ibldpc
