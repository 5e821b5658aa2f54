// ib_fast_cn.cu -- instantiations of ib_cn_fast_kernel<D, MATCH, EARLY> (see ib_kernels.cuh)
#include "kernel_tables.h"
namespace ibldpc {
template <bool EARLY>
NodeKernel cn_plain_sel(int d)
{
    switch (d) {
    case 2: return ib_cn_fast_kernel<2, false, EARLY, false>;
    case 3: return ib_cn_fast_kernel<3, false, EARLY, false>;
    case 4: return ib_cn_fast_kernel<4, false, EARLY, false>;
    case 5: return ib_cn_fast_kernel<5, false, EARLY, false>;
    case 6: return ib_cn_fast_kernel<6, false, EARLY, false>;
    case 7: return ib_cn_fast_kernel<7, false, EARLY, false>;
    case 8: return ib_cn_fast_kernel<8, false, EARLY, false>;
    case 9: return ib_cn_fast_kernel<9, false, EARLY, false>;
    case 10: return ib_cn_fast_kernel<10, false, EARLY, false>;
    default: return nullptr;
    }
}
template <bool EARLY>
NodeKernel cn_pair_sel(int d)
{
    switch (d) {
    case 4: return ib_cn_fast_kernel<4, false, EARLY, true>;
    case 5: return ib_cn_fast_kernel<5, false, EARLY, true>;
    case 6: return ib_cn_fast_kernel<6, false, EARLY, true>;
    case 7: return ib_cn_fast_kernel<7, false, EARLY, true>;
    case 8: return ib_cn_fast_kernel<8, false, EARLY, true>;
    case 9: return ib_cn_fast_kernel<9, false, EARLY, true>;
    case 10: return ib_cn_fast_kernel<10, false, EARLY, true>;
    default: return nullptr;
    }
}
// `match` = explicit matching look-up, needed by degree-2 checks only (all other degrees get the
// matching folded into their last-stage table at staging time, see stage_tables).
// `pair` selects the tail-pair variant (cn_word_pair), available for d >= 4.
NodeKernel cn_fast_kernel_for(int d, bool match, bool early, bool pair)
{
    if (match) {
        if (d != 2) return nullptr;
        return early ? (NodeKernel)ib_cn_fast_kernel<2, true, true, false> : (NodeKernel)ib_cn_fast_kernel<2, true, false, false>;
    }
    if (pair && d >= 4) return early ? cn_pair_sel<true>(d) : cn_pair_sel<false>(d);
    return early ? cn_plain_sel<true>(d) : cn_plain_sel<false>(d);
}
}  // namespace ibldpc
