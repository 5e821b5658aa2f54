// ib_t32_cn.cu -- instantiations of the |T| <= 32 check-node kernels ib_t32_kernel<kPhaseCn, EARLY, D>
#include "ib_kernels_t32.cuh"
namespace ibldpc {
template <bool EARLY>
static T32Kernel sel(int d)
{
    switch (d) {
    case 3: return ib_t32_kernel<kPhaseCn, EARLY, 3>;
    case 4: return ib_t32_kernel<kPhaseCn, EARLY, 4>;
    case 5: return ib_t32_kernel<kPhaseCn, EARLY, 5>;
    case 6: return ib_t32_kernel<kPhaseCn, EARLY, 6>;
    case 7: return ib_t32_kernel<kPhaseCn, EARLY, 7>;
    case 8: return ib_t32_kernel<kPhaseCn, EARLY, 8>;
    case 9: return ib_t32_kernel<kPhaseCn, EARLY, 9>;
    case 10: return ib_t32_kernel<kPhaseCn, EARLY, 10>;
    default: return nullptr;
    }
}
T32Kernel t32_cn_kernel(int d, bool early) { return early ? sel<true>(d) : sel<false>(d); }
}  // namespace ibldpc
