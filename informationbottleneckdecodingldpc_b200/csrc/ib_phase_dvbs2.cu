// ib_phase_dvbs2.cu -- instantiation of the fused per-phase kernels (ib_phase_n4.cuh) for one degree set
// (class = degree, or degree + 100 * (pair mode + 1): 2xx = tail pair, 1xx = plain chains)
#include "ib_phase_sets.h"
namespace ibldpc {
const PhaseSetOps* phase_ops_dvbs2()
{
    static const PhaseSetOps ops = make_phase_ops("dvbs2", DegreeSet<7, 106>{}, DegreeSet<8, 3, 2, 1>{});
    return &ops;
}
}  // namespace ibldpc
