// ib_n4_cn_v4.cu -- instantiations of ib_cn_n4_kernel<D, MATCH, EARLY, 4> (see ib_kernels_n4.cuh)
#include "kernel_tables.h"
#include "ib_kernels_n4.cuh"
namespace ibldpc {
template <bool EARLY>
static NodeKernel cn_n4_sel(int d)
{
    switch (d) {
    case 2: return ib_cn_n4_kernel<2, false, EARLY, 4, false>;
    case 3: return ib_cn_n4_kernel<3, false, EARLY, 4, false>;
    case 4: return ib_cn_n4_kernel<4, false, EARLY, 4, false>;
    case 5: return ib_cn_n4_kernel<5, false, EARLY, 4, false>;
    case 6: return ib_cn_n4_kernel<6, false, EARLY, 4, false>;
    case 7: return ib_cn_n4_kernel<7, false, EARLY, 4, false>;
    case 8: return ib_cn_n4_kernel<8, false, EARLY, 4, false>;
    case 9: return ib_cn_n4_kernel<9, false, EARLY, 4, false>;
    case 10: return ib_cn_n4_kernel<10, false, EARLY, 4, false>;
    default: return nullptr;
    }
}
// `match` = explicit matching look-up, needed by degree-2 checks only (all other degrees get the
// matching folded into their last-stage table at staging time, see stage_tables_n4).
NodeKernel cn_n4_kernel_v4(int d, bool match, bool early)
{
    if (match) {
        if (d != 2) return nullptr;
        return early ? (NodeKernel)ib_cn_n4_kernel<2, true, true, 4, false> : (NodeKernel)ib_cn_n4_kernel<2, true, false, 4, false>;
    }
    return early ? cn_n4_sel<true>(d) : cn_n4_sel<false>(d);
}
}  // namespace ibldpc
