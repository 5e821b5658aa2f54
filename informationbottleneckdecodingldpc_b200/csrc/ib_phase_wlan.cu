// ib_phase_wlan.cu -- instantiation of the fused per-phase kernels (ib_phase_n4.cuh) for one degree set
// (class = degree, or degree + 100 * (pair mode + 1): 2xx = tail pair, 1xx = plain chains)
#include "ib_phase_sets.h"
namespace ibldpc {
const PhaseSetOps* phase_ops_wlan()
{
    static const PhaseSetOps ops = make_phase_ops("wlan", DegreeSet<8, 7>{}, DegreeSet<11, 4, 3, 2>{});
    return &ops;
}
}  // namespace ibldpc
