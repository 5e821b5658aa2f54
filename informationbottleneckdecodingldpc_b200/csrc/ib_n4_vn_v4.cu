// ib_n4_vn_v4.cu -- instantiations of ib_vn_n4_kernel<D, 4> / ib_out_n4_kernel<D, 4> (see ib_kernels_n4.cuh)
// for D <= 6 (wider degrees spill with 4 words per lane and use the 2-word kernels)
#include "kernel_tables.h"
#include "ib_kernels_n4.cuh"
namespace ibldpc {
NodeKernel vn_n4_kernel_v4(int d, bool decide)
{
    if (!decide) {
        switch (d) {
        case 1: return ib_vn_n4_kernel<1, 4>;
        case 2: return ib_vn_n4_kernel<2, 4>;
        case 3: return ib_vn_n4_kernel<3, 4>;
        case 4: return ib_vn_n4_kernel<4, 4>;
        case 5: return ib_vn_n4_kernel<5, 4>;
        case 6: return ib_vn_n4_kernel<6, 4>;
        default: return nullptr;
        }
    }
    switch (d) {
    case 1: return ib_out_n4_kernel<1, 4>;
    case 2: return ib_out_n4_kernel<2, 4>;
    case 3: return ib_out_n4_kernel<3, 4>;
    case 4: return ib_out_n4_kernel<4, 4>;
    case 5: return ib_out_n4_kernel<5, 4>;
    case 6: return ib_out_n4_kernel<6, 4>;
    default: return nullptr;
    }
}
// 1024-thread CTAs for the update kernels of degree 2..4 (default for large batches: one table set and 32 warps per
// SM; measured on B200, C1 B=65536: 0.446 ms with 4 x 256 threads, 0.426 ms with 2 x 512, 0.403 ms with 1 x 1024)
NodeKernel vn_n4_kernel_v4_1024(int d)
{
    switch (d) {
    case 2: return ib_vn_n4_kernel<2, 4, 1024>;
    case 3: return ib_vn_n4_kernel<3, 4, 1024>;
    case 4: return ib_vn_n4_kernel<4, 4, 1024>;
    default: return nullptr;
    }
}
}  // namespace ibldpc
