// ib_n4_vn_pair.cu -- instantiations of the tail-pair variable-node kernels ib_vn_n4_pair_kernel<D, NT>
// (shift/LOP3 address arithmetic: measured on the DVB-S2-like code, the d_v = 8 kernel is 2 % slower with the dp4a form --
// its fma pipe already carries the look-up addresses and the nibble inserts; -DIBLDPC_DP4A_ALL for the A/B build)
#ifndef IBLDPC_DP4A_ALL
#define IBLDPC_NO_DP4A
#endif
#include "kernel_tables.h"
#include "ib_kernels_n4.cuh"
namespace ibldpc {
NodeKernel vn_n4_pair_kernel(int d, int threads)
{
    if (threads == 768) {   // one CTA per SM, 24 warps at <= 85 registers (degrees 8-9 need 116-124 unconstrained, spill at 64)
        switch (d) {
        case 8: return ib_vn_n4_pair_kernel<8, 768>;
        case 9: return ib_vn_n4_pair_kernel<9, 768>;
        default: return nullptr;
        }
    }
    if (threads == 512) {
        switch (d) {
        case 3: return ib_vn_n4_pair_kernel<3, 512>;
        case 4: return ib_vn_n4_pair_kernel<4, 512>;
        case 5: return ib_vn_n4_pair_kernel<5, 512>;
        case 6: return ib_vn_n4_pair_kernel<6, 512>;
        case 7: return ib_vn_n4_pair_kernel<7, 512>;
        case 8: return ib_vn_n4_pair_kernel<8, 512>;
        case 9: return ib_vn_n4_pair_kernel<9, 512>;
        case 10: return ib_vn_n4_pair_kernel<10, 512>;
        case 11: return ib_vn_n4_pair_kernel<11, 512>;
        case 12: return ib_vn_n4_pair_kernel<12, 512>;
        default: return nullptr;
        }
    }
    switch (d) {
    case 3: return ib_vn_n4_pair_kernel<3, 256>;
    case 4: return ib_vn_n4_pair_kernel<4, 256>;
    case 5: return ib_vn_n4_pair_kernel<5, 256>;
    case 6: return ib_vn_n4_pair_kernel<6, 256>;
    case 7: return ib_vn_n4_pair_kernel<7, 256>;
    case 8: return ib_vn_n4_pair_kernel<8, 256>;
    case 9: return ib_vn_n4_pair_kernel<9, 256>;
    case 10: return ib_vn_n4_pair_kernel<10, 256>;
    case 11: return ib_vn_n4_pair_kernel<11, 256>;
    case 12: return ib_vn_n4_pair_kernel<12, 256>;
    default: return nullptr;
    }
}
}  // namespace ibldpc
