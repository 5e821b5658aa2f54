// ib_n4_cn_pair.cu -- instantiations of the tail-pair check-node kernels
// ib_cn_n4_kernel<D, false, EARLY, 2, true, NT> (512 threads per CTA up to degree 8, 256 above)
#include "kernel_tables.h"
#include "ib_kernels_n4.cuh"
namespace ibldpc {
template <bool EARLY>
static NodeKernel cn_n4_pair_sel(int d)
{
    switch (d) {
    case 4: return ib_cn_n4_kernel<4, false, EARLY, 2, true, 512>;
    case 5: return ib_cn_n4_kernel<5, false, EARLY, 2, true, 512>;
    case 6: return ib_cn_n4_kernel<6, false, EARLY, 2, true, 512>;
    case 7: return ib_cn_n4_kernel<7, false, EARLY, 2, true, 512>;
    case 8: return ib_cn_n4_kernel<8, false, EARLY, 2, true, 512>;
    case 9: return ib_cn_n4_kernel<9, false, EARLY, 2, true, 256>;
    case 10: return ib_cn_n4_kernel<10, false, EARLY, 2, true, 256>;
    default: return nullptr;
    }
}
int cn_n4_pair_threads(int d) { return d <= 8 ? 512 : 256; }
// 1024-thread CTAs (default up to degree 8: one 64-96 KB table set and 32 warps per SM; measured on B200: C1 check-node
// phase 0.512 -> 0.488 ms, 802.11n 0.216 -> 0.192 ms against 2 x 512 threads; IBLDPC_CN_THREADS=512 selects the latter)
template <bool EARLY>
static NodeKernel cn_n4_pair_sel_1024(int d)
{
    switch (d) {
    case 4: return ib_cn_n4_kernel<4, false, EARLY, 2, true, 1024>;
    case 5: return ib_cn_n4_kernel<5, false, EARLY, 2, true, 1024>;
    case 6: return ib_cn_n4_kernel<6, false, EARLY, 2, true, 1024>;
    case 7: return ib_cn_n4_kernel<7, false, EARLY, 2, true, 1024>;
    case 8: return ib_cn_n4_kernel<8, false, EARLY, 2, true, 1024>;
    default: return nullptr;
    }
}
NodeKernel cn_n4_pair_kernel_1024(int d, bool early) { return early ? cn_n4_pair_sel_1024<true>(d) : cn_n4_pair_sel_1024<false>(d); }
NodeKernel cn_n4_pair_kernel(int d, bool early) { return early ? cn_n4_pair_sel<true>(d) : cn_n4_pair_sel<false>(d); }
}  // namespace ibldpc
