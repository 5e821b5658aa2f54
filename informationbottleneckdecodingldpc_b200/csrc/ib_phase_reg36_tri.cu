// ib_phase_reg36_tri.cu -- fused per-phase kernels of the regular (3,6) set through the three-input tables
// (class = degree + 300: ib_triple_n4.cuh; image = [tail-pair rows][F, 128 KB][stage columns])
#include "ib_phase_sets.h"
namespace ibldpc {
const PhaseSetOps* phase_ops_reg36_tri()
{
    static const PhaseSetOps ops = make_phase_ops("reg36_tri", DegreeSet<306>{}, DegreeSet<303>{});
    return &ops;
}
}  // namespace ibldpc
