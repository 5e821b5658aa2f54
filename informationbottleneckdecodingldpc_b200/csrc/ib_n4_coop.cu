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

namespace {
template <int... Ds>
bool same_set(DegreeSet<Ds...>, const int* deg, int n)
{
    constexpr int want[] = {Ds...};
    if (n != (int)sizeof...(Ds)) return false;
    for (int i = 0; i < n; ++i) {
        bool found = false;
        for (int w : want) found |= (w == deg[i]);
        if (!found) return false;
    }
    return true;   // the class lists hold distinct degrees, so equal size + containment = equality
}
template <typename CnSet, typename VnSet>
CoopMultiKernel pick(const int* cn_deg, int n_cn, const int* vn_deg, int n_vn, bool early, int T, bool match, int* smem_bytes)
{
    if (!same_set(CnSet{}, cn_deg, n_cn) || !same_set(VnSet{}, vn_deg, n_vn)) return nullptr;
    *smem_bytes = coop_multi_smem_bytes(CnSet{}, VnSet{}, T, match);
    return early ? (CoopMultiKernel)ib_decode_coop_multi_kernel<CnSet, VnSet, true>
                 : (CoopMultiKernel)ib_decode_coop_multi_kernel<CnSet, VnSet, false>;
}
}  // namespace

CoopMultiKernel coop_multi_kernel_for(const int* cn_deg, int n_cn, const int* vn_deg, int n_vn, bool early, int T, bool match,
                                      int* smem_bytes)
{
    // IEEE 802.11n rate 1/2 (generate_802.11_matrix.py) and DVB-S2 rate 1/2 (DVB-S2/decoder_config_generation.py:32-34)
    if (auto k = pick<DegreeSet<7, 8>, DegreeSet<2, 3, 4, 11>>(cn_deg, n_cn, vn_deg, n_vn, early, T, match, smem_bytes)) return k;
    if (auto k = pick<DegreeSet<6, 7>, DegreeSet<1, 2, 3, 8>>(cn_deg, n_cn, vn_deg, n_vn, early, T, match, smem_bytes)) return k;
    return nullptr;
}
}  // namespace ibldpc
