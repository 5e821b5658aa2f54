// ib_n4_vn3.cu -- instantiations of the three-input-table kernels (ib_triple_n4.cuh): degree-3 variable nodes, checks of degree 6..8
#include "kernel_tables.h"
#include "ib_triple_n4.cuh"
namespace ibldpc {
NodeKernel vn3_n4_kernel(int vec) { return vec == 4 ? ib_vn3_n4_kernel<4, 1024> : ib_vn3_n4_kernel<2, 1024>; }
template <bool EARLY>
static NodeKernel cn_n4_tri_sel(int d)
{
    switch (d) {
    case 6: return ib_cn_n4_tri_kernel<6, EARLY, 1024>;
    case 7: return ib_cn_n4_tri_kernel<7, EARLY, 1024>;
    case 8: return ib_cn_n4_tri_kernel<8, EARLY, 1024>;
    default: return nullptr;
    }
}
NodeKernel cn_n4_tri_kernel(int d, bool early) { return early ? cn_n4_tri_sel<true>(d) : cn_n4_tri_sel<false>(d); }
}  // namespace ibldpc
