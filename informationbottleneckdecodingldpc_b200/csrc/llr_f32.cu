// llr_f32.cu -- instantiations of the min-sum / BP kernels for float messages (see llr_kernels.cuh)
#include "kernel_tables.h"
namespace ibldpc {
namespace {
template <int ALGO>
LlrNodeKernel cn_sel(int d)
{
    switch (d) {
    case 2: return llr_cn_kernel<float, ALGO, 2>;
    case 3: return llr_cn_kernel<float, ALGO, 3>;
    case 4: return llr_cn_kernel<float, ALGO, 4>;
    case 5: return llr_cn_kernel<float, ALGO, 5>;
    case 6: return llr_cn_kernel<float, ALGO, 6>;
    case 7: return llr_cn_kernel<float, ALGO, 7>;
    case 8: return llr_cn_kernel<float, ALGO, 8>;
    case 9: return llr_cn_kernel<float, ALGO, 9>;
    case 10: return llr_cn_kernel<float, ALGO, 10>;
    default: return llr_cn_kernel<float, ALGO, 0>;
    }
}
template <int MODE>
LlrNodeKernel vn_sel(int d)
{
    switch (d) {
    case 1: return llr_vn_kernel<float, MODE, 1>;
    case 2: return llr_vn_kernel<float, MODE, 2>;
    case 3: return llr_vn_kernel<float, MODE, 3>;
    case 4: return llr_vn_kernel<float, MODE, 4>;
    case 5: return llr_vn_kernel<float, MODE, 5>;
    case 6: return llr_vn_kernel<float, MODE, 6>;
    case 7: return llr_vn_kernel<float, MODE, 7>;
    case 8: return llr_vn_kernel<float, MODE, 8>;
    case 9: return llr_vn_kernel<float, MODE, 9>;
    case 10: return llr_vn_kernel<float, MODE, 10>;
    case 11: return llr_vn_kernel<float, MODE, 11>;
    case 12: return llr_vn_kernel<float, MODE, 12>;
    default: return llr_vn_kernel<float, MODE, 0>;
    }
}
}  // namespace
LlrNodeKernel llr_cn_kernel_f32(int algo, int d) { return algo == 0 ? cn_sel<0>(d) : cn_sel<1>(d); }
LlrNodeKernel llr_vn_kernel_f32(int mode, int d) { return mode == 0 ? vn_sel<0>(d) : mode == 1 ? vn_sel<1>(d) : vn_sel<2>(d); }
LlrSynKernel llr_syndrome_kernel_f32() { return llr_syndrome_kernel<float>; }
LlrShflKernel llr_cn_minsum_shfl_kernel_f32() { return llr_cn_minsum_shfl_kernel<float>; }
}  // namespace ibldpc
