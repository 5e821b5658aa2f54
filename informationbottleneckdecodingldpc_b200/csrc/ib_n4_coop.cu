// ib_n4_coop.cu -- instantiations of ib_decode_coop_kernel<DC, DV, EARLY> (see ib_coop_n4.cuh) for the regular
// (d_v, d_c) ensembles of the reference's drivers and tests
#include "ib_coop_n4.cuh"
namespace ibldpc {
#define COOP_CASE(DC_, DV_)                                                                                        \
    if (dc == DC_ && dv == DV_) {                                                                                  \
        *smem_bytes = coop_smem_bytes<DC_, DV_>(T, match);                                                         \
        return early ? (CoopKernel)ib_decode_coop_kernel<DC_, DV_, true> : (CoopKernel)ib_decode_coop_kernel<DC_, DV_, false>; \
    }
CoopKernel coop_kernel_for(int dc, int dv, bool early, int T, bool match, int* smem_bytes)
{
    COOP_CASE(6, 3)
    COOP_CASE(5, 3)
    COOP_CASE(4, 3)
    COOP_CASE(4, 2)
    COOP_CASE(8, 4)
    return nullptr;
}
#undef COOP_CASE
}  // namespace ibldpc
