// kernel_tables.h -- per-degree kernel instantiations live in their own translation units
// (ib_fast_cn.cu, ib_fast_vn.cu, llr_f32.cu, llr_f64.cu) so that they compile in parallel;
// the host code picks them through these selectors.
#pragma once
#include "ib_kernels.cuh"
#include "llr_kernels.cuh"

namespace ibldpc {
using NodeKernel = void (*)(IbArgs, const int*, int);
using LlrNodeKernel = void (*)(LlrArgs, const int*, int);
using LlrSynKernel = void (*)(LlrArgs);
using LlrShflKernel = void (*)(LlrArgs, int);

NodeKernel cn_fast_kernel_for(int d, bool match, bool early, bool pair);
NodeKernel vn_fast_kernel_for(int d, bool decide, bool match);
// packed-nibble fast path (ib_kernels_n4.cuh); v2 / v4 = 32-bit words per lane and message
// (check nodes: 2; variable nodes: 4 up to degree 6, else 2)
NodeKernel cn_n4_kernel_v2(int d, bool match, bool early);
NodeKernel cn_n4_pair_kernel(int d, bool early);   // tail-pair variant (2 words per lane), d >= 4
int cn_n4_pair_threads(int d);
NodeKernel cn_n4_pair_kernel_1024(int d, bool early);   // d <= 8 (default); cn_n4_pair_kernel: 512 threads (d <= 8) / 256                     // threads per CTA of that kernel
NodeKernel vn_n4_pair_kernel(int d, int threads);   // tail-pair variable-node update, d >= 3, threads = 256 / 512
NodeKernel vn_n4_kernel_v2(int d, bool decide);
NodeKernel vn_n4_kernel_v4(int d, bool decide);
NodeKernel cn_n4_tri_kernel(int d, bool early);  // check nodes of degree 6..8: three-input table of the first two stages + tail-pair rows
NodeKernel vn3_n4_kernel(int vec);         // degree-3 update through the three-input table (ib_triple_n4.cuh), 1024-thread CTAs
NodeKernel vn_n4_kernel_v4_1024(int d);   // update kernels of degree 2..4 in 1024-thread CTAs (default)
LlrNodeKernel llr_cn_kernel_for(bool f64, int algo, int d);
LlrNodeKernel llr_vn_kernel_for(bool f64, int mode, int d);
LlrSynKernel llr_syndrome_kernel_for(bool f64);
LlrNodeKernel llr_cn_kernel_f32(int algo, int d);
LlrNodeKernel llr_vn_kernel_f32(int mode, int d);
LlrNodeKernel llr_cn_kernel_f64(int algo, int d);
LlrNodeKernel llr_vn_kernel_f64(int mode, int d);
LlrSynKernel llr_syndrome_kernel_f32();
LlrSynKernel llr_syndrome_kernel_f64();
LlrShflKernel llr_cn_minsum_shfl_kernel_for(bool f64);
LlrShflKernel llr_cn_minsum_shfl_kernel_f32();
LlrShflKernel llr_cn_minsum_shfl_kernel_f64();
}  // namespace ibldpc
