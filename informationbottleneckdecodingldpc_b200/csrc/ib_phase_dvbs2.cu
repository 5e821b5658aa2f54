// ib_phase_dvbs2.cu -- instantiation of the fused per-phase kernels (ib_phase_n4.cuh) for one degree set
#include "ib_phase_sets.h"
namespace ibldpc {
const PhaseSetOps* phase_ops_dvbs2()
{
    static const PhaseSetOps ops = make_phase_ops("dvbs2", DegreeSet<7, 6>{}, DegreeSet<8, 3, 2, 1>{});
    return &ops;
}
}  // namespace ibldpc
