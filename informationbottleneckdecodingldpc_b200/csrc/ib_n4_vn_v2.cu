// ib_n4_vn_v2.cu -- instantiations of ib_vn_n4_kernel<D, 2> / ib_out_n4_kernel<D, 2> (see ib_kernels_n4.cuh)
#include "kernel_tables.h"
#include "ib_kernels_n4.cuh"
namespace ibldpc {
NodeKernel vn_n4_kernel_v2(int d, bool decide)
{
    if (!decide) {
        switch (d) {
        case 1: return ib_vn_n4_kernel<1, 2>;
        case 2: return ib_vn_n4_kernel<2, 2>;
        case 3: return ib_vn_n4_kernel<3, 2>;
        case 4: return ib_vn_n4_kernel<4, 2>;
        case 5: return ib_vn_n4_kernel<5, 2>;
        case 6: return ib_vn_n4_kernel<6, 2>;
        case 7: return ib_vn_n4_kernel<7, 2>;
        case 8: return ib_vn_n4_kernel<8, 2>;
        case 9: return ib_vn_n4_kernel<9, 2>;
        case 10: return ib_vn_n4_kernel<10, 2>;
        case 11: return ib_vn_n4_kernel<11, 2>;
        case 12: return ib_vn_n4_kernel<12, 2>;
        default: return nullptr;
        }
    }
    switch (d) {
    case 1: return ib_out_n4_kernel<1, 2>;
    case 2: return ib_out_n4_kernel<2, 2>;
    case 3: return ib_out_n4_kernel<3, 2>;
    case 4: return ib_out_n4_kernel<4, 2>;
    case 5: return ib_out_n4_kernel<5, 2>;
    case 6: return ib_out_n4_kernel<6, 2>;
    case 7: return ib_out_n4_kernel<7, 2>;
    case 8: return ib_out_n4_kernel<8, 2>;
    case 9: return ib_out_n4_kernel<9, 2>;
    case 10: return ib_out_n4_kernel<10, 2>;
    case 11: return ib_out_n4_kernel<11, 2>;
    case 12: return ib_out_n4_kernel<12, 2>;
    default: return nullptr;
    }
}
}  // namespace ibldpc
