// ib_t32_vn.cu -- instantiations of ib_t32_kernel<kPhaseVn, false, D> (|T| <= 32 family, ib_kernels_t32.cuh)
#include "ib_kernels_t32.cuh"
namespace ibldpc {
T32Kernel t32_vn_kernel(int d)
{
    switch (d) {
    case 1: return ib_t32_kernel<kPhaseVn, false, 1>;
    case 2: return ib_t32_kernel<kPhaseVn, false, 2>;
    case 3: return ib_t32_kernel<kPhaseVn, false, 3>;
    case 4: return ib_t32_kernel<kPhaseVn, false, 4>;
    case 5: return ib_t32_kernel<kPhaseVn, false, 5>;
    case 6: return ib_t32_kernel<kPhaseVn, false, 6>;
    case 7: return ib_t32_kernel<kPhaseVn, false, 7>;
    case 8: return ib_t32_kernel<kPhaseVn, false, 8>;
    case 9: return ib_t32_kernel<kPhaseVn, false, 9>;
    case 10: return ib_t32_kernel<kPhaseVn, false, 10>;
    case 11: return ib_t32_kernel<kPhaseVn, false, 11>;
    case 12: return ib_t32_kernel<kPhaseVn, false, 12>;
    default: return nullptr;
    }
}
}  // namespace ibldpc
