// llr_f64.cu -- instantiations of the min-sum / BP kernels for double messages (see llr_kernels.cuh)
#include "kernel_tables.h"
namespace ibldpc {
namespace {
template <int ALGO>
LlrNodeKernel cn_sel(int d)
{
    switch (d) {
    case 2: return llr_cn_kernel<double, ALGO, 2>;
    case 3: return llr_cn_kernel<double, ALGO, 3>;
    case 4: return llr_cn_kernel<double, ALGO, 4>;
    case 5: return llr_cn_kernel<double, ALGO, 5>;
    case 6: return llr_cn_kernel<double, ALGO, 6>;
    case 7: return llr_cn_kernel<double, ALGO, 7>;
    case 8: return llr_cn_kernel<double, ALGO, 8>;
    case 9: return llr_cn_kernel<double, ALGO, 9>;
    case 10: return llr_cn_kernel<double, ALGO, 10>;
    default: return llr_cn_kernel<double, ALGO, 0>;
    }
}
template <int MODE>
LlrNodeKernel vn_sel(int d)
{
    switch (d) {
    case 1: return llr_vn_kernel<double, MODE, 1>;
    case 2: return llr_vn_kernel<double, MODE, 2>;
    case 3: return llr_vn_kernel<double, MODE, 3>;
    case 4: return llr_vn_kernel<double, MODE, 4>;
    case 5: return llr_vn_kernel<double, MODE, 5>;
    case 6: return llr_vn_kernel<double, MODE, 6>;
    case 7: return llr_vn_kernel<double, MODE, 7>;
    case 8: return llr_vn_kernel<double, MODE, 8>;
    case 9: return llr_vn_kernel<double, MODE, 9>;
    case 10: return llr_vn_kernel<double, MODE, 10>;
    case 11: return llr_vn_kernel<double, MODE, 11>;
    case 12: return llr_vn_kernel<double, MODE, 12>;
    default: return llr_vn_kernel<double, MODE, 0>;
    }
}
}  // namespace
LlrNodeKernel llr_cn_kernel_f64(int algo, int d) { return algo == 0 ? cn_sel<0>(d) : algo == 3 ? cn_sel<3>(d) : algo == 2 ? cn_sel<2>(d) : cn_sel<1>(d); }
LlrNodeKernel llr_vn_kernel_f64(int mode, int d) { return mode == 0 ? vn_sel<0>(d) : mode == 1 ? vn_sel<1>(d) : vn_sel<2>(d); }
LlrSynKernel llr_syndrome_kernel_f64() { return llr_syndrome_kernel<double>; }
LlrShflKernel llr_cn_minsum_shfl_kernel_f64() { return llr_cn_minsum_shfl_kernel<double>; }
}  // namespace ibldpc

namespace ibldpc {
LlrNodeKernel llr_cn_kernel_for(bool f64, int algo, int d) { return f64 ? llr_cn_kernel_f64(algo, d) : llr_cn_kernel_f32(algo, d); }
LlrNodeKernel llr_vn_kernel_for(bool f64, int mode, int d) { return f64 ? llr_vn_kernel_f64(mode, d) : llr_vn_kernel_f32(mode, d); }
LlrSynKernel llr_syndrome_kernel_for(bool f64) { return f64 ? llr_syndrome_kernel_f64() : llr_syndrome_kernel_f32(); }
LlrShflKernel llr_cn_minsum_shfl_kernel_for(bool f64) { return f64 ? llr_cn_minsum_shfl_kernel_f64() : llr_cn_minsum_shfl_kernel_f32(); }
}  // namespace ibldpc
