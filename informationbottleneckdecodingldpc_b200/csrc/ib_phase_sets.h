// ib_phase_sets.h -- degree sets the fused per-phase kernels (ib_phase_n4.cuh) are instantiated for; one translation
// unit per set so that they compile in parallel.  Classes are listed heaviest (highest degree) first.
#pragma once
#include <vector>

#include "ib_phase_n4.cuh"

namespace ibldpc {

struct PhaseSetOps {
    const char* name;
    std::vector<int> cn_deg, vn_deg;            // degree sets, heaviest first
    PhaseLayoutRt cn_layout, vn_layout, out_layout;
    PhaseKernel cn_kernel[2];                   // [early]
    PhaseKernel vn_kernel, out_kernel;
    PhaseKernel cn_pf_kernel[2], vn_pf_kernel;  // per-frame early termination (ib_perframe.cu); [0] syndrome flags in shared memory, [1] global
    PhaseDecideKernel pf_decide_kernel;         // deferred decision of the converged frames
    CoopPhaseKernel coop_kernel[2];             // [early] whole decode in one cooperative launch, B <= kLaneModeMaxFrames
    int coop_threads;
};

constexpr int kCoopPhaseThreads = 512;          // 128 registers per thread: the degree-11 lane bodies keep their rows and words in registers

template <int... Cs, int... Vs>
PhaseSetOps make_phase_ops(const char* name, DegreeSet<Cs...> cs, DegreeSet<Vs...> vs)
{
    PhaseSetOps o;
    o.name = name;
    o.cn_deg = {spec_deg(Cs)...};
    o.vn_deg = {spec_deg(Vs)...};
    o.cn_layout = phase_layout_rt<kPhaseCn>(cs);
    o.vn_layout = phase_layout_rt<kPhaseVn>(vs);
    o.out_layout = phase_layout_rt<kPhaseOut>(vs);
    o.cn_kernel[0] = ib_phase_kernel<kPhaseCn, false, Cs...>;
    o.cn_kernel[1] = ib_phase_kernel<kPhaseCn, true, Cs...>;
    o.vn_kernel = ib_phase_kernel<kPhaseVn, false, Vs...>;
    o.out_kernel = ib_phase_kernel<kPhaseOut, false, Vs...>;
    o.cn_pf_kernel[0] = ib_phase_pf_kernel<kPhaseCn, 1, Cs...>;
    o.cn_pf_kernel[1] = ib_phase_pf_kernel<kPhaseCn, 2, Cs...>;
    o.vn_pf_kernel = ib_phase_pf_kernel<kPhaseVn, 1, Vs...>;
    o.pf_decide_kernel = ib_phase_pfdecide_kernel<Vs...>;
    o.coop_kernel[0] = ib_coop_phase_kernel<kCoopPhaseThreads, false, DegreeSet<Cs...>, DegreeSet<Vs...>>;
    o.coop_kernel[1] = ib_coop_phase_kernel<kCoopPhaseThreads, true, DegreeSet<Cs...>, DegreeSet<Vs...>>;
    o.coop_threads = kCoopPhaseThreads;
    return o;
}

const PhaseSetOps* phase_ops_wlan();     // IEEE 802.11n rate 1/2: d_c {8,7}, d_v {11,4,3,2}   (generate_802.11_matrix.py)
const PhaseSetOps* phase_ops_dvbs2();    // DVB-S2 rate 1/2: d_c {7,6}, d_v {8,3,2,1}           (DVB-S2/decoder_config_generation.py:32-34)
const PhaseSetOps* phase_ops_reg36_tri();  // the same code through the three-input tables (d + 300): per-frame early termination
const PhaseSetOps* phase_ops_reg36();    // regular (3,6): d_c {6}, d_v {3}                       (Regular_LDPC_Decoding/BPSK)

}  // namespace ibldpc
