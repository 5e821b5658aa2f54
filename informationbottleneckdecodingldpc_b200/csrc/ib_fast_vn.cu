// ib_fast_vn.cu -- instantiations of ib_vn_fast_kernel<D, MATCH> / ib_out_fast_kernel<D> (see ib_kernels.cuh)
#include "kernel_tables.h"
namespace ibldpc {
template <bool MATCH>
NodeKernel vn_fast_kernel_sel(int d)
{
#define VNK(D) case D: return ib_vn_fast_kernel<D, MATCH>;
    switch (d) {
        VNK(1) VNK(2) VNK(3) VNK(4) VNK(5) VNK(6) VNK(7) VNK(8) VNK(9) VNK(10) VNK(11) VNK(12)
    default: return nullptr;
    }
#undef VNK
}
NodeKernel vn_fast_kernel_for(int d, bool decide, bool match)
{
    (void)match;   // message alignment is folded into the staged tables: one instantiation serves both
    if (!decide) return vn_fast_kernel_sel<false>(d);
#define VNK(D) case D: return ib_out_fast_kernel<D>;
    switch (d) {
        VNK(1) VNK(2) VNK(3) VNK(4) VNK(5) VNK(6) VNK(7) VNK(8) VNK(9) VNK(10) VNK(11) VNK(12)
    default: return nullptr;
    }
#undef VNK
}

}  // namespace ibldpc
