// ib_n4_vn3.cu -- instantiations of the three-input-table kernels (ib_triple_n4.cuh): degree-3 variable nodes, degree-6 checks
#include "kernel_tables.h"
#include "ib_triple_n4.cuh"
namespace ibldpc {
NodeKernel vn3_n4_kernel(int vec) { return vec == 4 ? ib_vn3_n4_kernel<4, 1024> : ib_vn3_n4_kernel<2, 1024>; }
NodeKernel cn6_n4_tri_kernel(bool early) { return early ? ib_cn6_n4_tri_kernel<true, 1024> : ib_cn6_n4_tri_kernel<false, 1024>; }
}  // namespace ibldpc
