// ib_fast_cn.cu -- instantiations of ib_cn_fast_kernel<D, MATCH, EARLY> (see ib_kernels.cuh)
#include "kernel_tables.h"
namespace ibldpc {
template <bool MATCH, bool EARLY>
NodeKernel cn_fast_kernel_sel(int d)
{
    switch (d) {
    case 2: return ib_cn_fast_kernel<2, MATCH, EARLY>;
    case 3: return ib_cn_fast_kernel<3, MATCH, EARLY>;
    case 4: return ib_cn_fast_kernel<4, MATCH, EARLY>;
    case 5: return ib_cn_fast_kernel<5, MATCH, EARLY>;
    case 6: return ib_cn_fast_kernel<6, MATCH, EARLY>;
    case 7: return ib_cn_fast_kernel<7, MATCH, EARLY>;
    case 8: return ib_cn_fast_kernel<8, MATCH, EARLY>;
    case 9: return ib_cn_fast_kernel<9, MATCH, EARLY>;
    case 10: return ib_cn_fast_kernel<10, MATCH, EARLY>;
    default: return nullptr;
    }
}
// `match` = explicit matching look-up, needed by degree-2 checks only (all other degrees get the
// matching folded into their last-stage table at staging time, see stage_tables).
NodeKernel cn_fast_kernel_for(int d, bool match, bool early)
{
    if (match) {
        if (d != 2) return nullptr;
        return early ? (NodeKernel)ib_cn_fast_kernel<2, true, true> : (NodeKernel)ib_cn_fast_kernel<2, true, false>;
    }
    return early ? cn_fast_kernel_sel<false, true>(d) : cn_fast_kernel_sel<false, false>(d);
}
}  // namespace ibldpc
