// ib_phase_reg36.cu -- instantiation of the fused per-phase kernels (ib_phase_n4.cuh) for one degree set
// (class = degree, or degree + 100 * (pair mode + 1): 2xx = tail pair, 1xx = plain chains)
#include "ib_phase_sets.h"
namespace ibldpc {
const PhaseSetOps* phase_ops_reg36()
{
    static const PhaseSetOps ops = make_phase_ops("reg36", DegreeSet<6>{}, DegreeSet<3>{});
    return &ops;
}
}  // namespace ibldpc
