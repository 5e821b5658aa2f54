// ib_coop_n4.cuh -- whole-decode cooperative kernel for small batches of REGULAR codes (one check-node and one
// variable-node degree), packed-nibble family.
//
// A decode is 2 * i_max dependent phases.  Launched one kernel per phase, a batch below ~1000 frames pays the fixed
// cost of every launch (launch latency + table staging + one node per warp, ~17 us) 100 times: 1.75 ms per decode of
// the (3,6) n=8000 code for any B <= 512.  Here the whole flooding schedule
//   send + checknode_update_iter0 -> { varnode_update(it) -> checknode_update(it) + calc_syndrome } -> calc_varnode_output
// (discrete_LDPC_decoder.py:202-295) runs in ONE cooperative launch with a grid-wide barrier between the phases; the
// phase bodies are the device functions of the per-phase kernels (cn_loop_n4 / vn_loop_n4), so results, the
// batch-granular stop rule and i_num are identical.
#pragma once
#include <cooperative_groups.h>

#include "ib_kernels_n4.cuh"

namespace ibldpc {

struct CoopArgs {
    const int* cn_nodes;     // node lists of the single check / variable class
    const int* vn_nodes;
    int n_cn, n_vn;
    const uint8_t* cn8;      // whole tables, reference order
    const uint8_t* vn8;
    const uint8_t* mc8;      // matching vectors or nullptr
    const uint8_t* mv8;
    const uint8_t* cn_pair;  // composed tail-pair rows, one block per iteration block
    int DCmax, DVmax;        // CN_DEGREE / VN_DEGREE of the tables
};

constexpr int kCoopThreads = 512;

template <int DC, int DV>
constexpr int coop_smem_bytes(int T, bool match)
{
    constexpr bool PAIR = DC >= 6;
    const int cn = (PAIR ? (int)kPairBytes : 0) + n4_table_bytes(n4_cn_words(DC, false)) + stage_scratch_bytes(DC - 2, T, match ? DC : 0);
    const int vn = n4_table_bytes(n4_vn_words(DV, false)) + stage_scratch_bytes(DV - 1, T, match ? DV : 0);
    const int ou = n4_table_bytes(n4_vn_words(DV, true)) + stage_scratch_bytes(DV, T, 0);
    return cn > vn ? (cn > ou ? cn : ou) : (vn > ou ? vn : ou);
}

template <int DC, int DV, bool EARLY>
__global__ void __launch_bounds__(kCoopThreads, 2) ib_decode_coop_kernel(IbArgs a, CoopArgs c)
{
    namespace cg = cooperative_groups;
    static_assert(DC >= 3 && DV >= 2, "degree-2 checks / degree-1 variable nodes use the per-phase kernels");
    constexpr bool PAIR = DC >= 6;
    constexpr int NT = kCoopThreads, VEC = 2;
    extern __shared__ __align__(16) uint32_t s_all[];
    cg::grid_group grid = cg::this_grid();
    const int T = a.T, TT = T * T;
    const uint8_t* ptab = reinterpret_cast<const uint8_t*>(s_all);
    uint32_t* s_cn = s_all + (PAIR ? kPairBytes / 4 : 0);

    auto cn_phase = [&](int it) {
        IbArgs b = a;
        const int blk = it + 1;                 // table block: 0 = iteration-0 tables
        b.it = it; b.iter0 = (it < 0);
        b.lut = c.cn8 + (size_t)blk * (c.DCmax - 2) * TT;
        b.match = c.mc8 ? c.mc8 + (size_t)blk * c.DCmax * T : nullptr;
        b.nst = DC - 2; b.dmax_match = c.mc8 ? DC : 0;
        b.xp_col = PAIR ? DC - 5 : -1;
        if (PAIR) {
            const uint2* src = reinterpret_cast<const uint2*>(c.cn_pair + (size_t)blk * TT * 8);
            uint2* dst = reinterpret_cast<uint2*>(s_all);
            for (int i = threadIdx.x; i < kTS * kTS * kPairSlots; i += NT) {
                const int r = i / kPairSlots, ra = r / kTS, rb = r - ra * kTS;
                dst[i] = (ra < T && rb < T) ? src[ra * T + rb] : make_uint2(0u, 0u);
            }
        }
        stage_tables_n4<n4_cn_words(DC, false), NT>(s_cn, b, b.lut);
        __syncthreads();
        const uint32_t syn = a.B <= kLaneModeMaxFrames
                                 ? cn_lanes_n4<DC, false, EARLY, PAIR, NT>(b, reinterpret_cast<const uint8_t*>(s_cn), ptab, c.cn_nodes, c.n_cn)
                                 : cn_loop_n4<DC, false, EARLY, VEC, PAIR, NT>(b, reinterpret_cast<const uint8_t*>(s_cn), ptab, c.cn_nodes, c.n_cn);
        if (EARLY && it >= 0) {
            const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
            if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[it], 1);
        }
    };
    auto vn_phase = [&](int it) {
        IbArgs b = a;
        b.it = it; b.iter0 = 0;
        b.lut = c.vn8 + (size_t)it * c.DVmax * TT;
        b.match = c.mv8 ? c.mv8 + (size_t)it * c.DVmax * T : nullptr;
        b.nst = DV - 1; b.dmax_match = c.mv8 ? DV : 0;
        b.xp_col = -1;
        stage_tables_n4<n4_vn_words(DV, false), NT>(s_all, b, b.lut);
        __syncthreads();
        if (a.B <= kLaneModeMaxFrames) vn_lanes_n4<DV, false, NT>(b, reinterpret_cast<const uint8_t*>(s_all), c.vn_nodes, c.n_vn);
        else vn_loop_n4<DV, false, VEC, false, NT>(b, reinterpret_cast<const uint8_t*>(s_all), nullptr, c.vn_nodes, c.n_vn);
    };

    cn_phase(-1);
    grid.sync();
    int passes = 0;
    for (int it = 0; it < a.imax - 1; ++it) {
        // reference stop rule (discrete_LDPC_decoder.py:233-276): pass `it` runs iff it == 0 or pass it-1 left a
        // non-zero syndrome somewhere in the batch; flags[] were written before the last grid barrier
        if (EARLY && a.early && it >= 1 && *reinterpret_cast<volatile int*>(&a.flags[it - 1]) == 0) break;
        vn_phase(it);
        grid.sync();
        cn_phase(it);
        grid.sync();
        passes = it + 1;
    }
    {   // calc_varnode_output with the VN table of iteration i_num - 1
        if (blockIdx.x == 0 && threadIdx.x == 0) *a.inum = passes + 1;
        IbArgs b = a;
        b.it = passes; b.iter0 = 0;
        b.lut = c.vn8 + (size_t)passes * c.DVmax * TT;
        b.match = nullptr; b.nst = DV; b.dmax_match = 0; b.xp_col = -1;
        stage_tables_n4<n4_vn_words(DV, true), NT>(s_all, b, b.lut);
        __syncthreads();
        if (a.B <= kLaneModeMaxFrames) vn_lanes_n4<DV, true, NT>(b, reinterpret_cast<const uint8_t*>(s_all), c.vn_nodes, c.n_vn);
        else vn_loop_n4<DV, true, VEC, false, NT>(b, reinterpret_cast<const uint8_t*>(s_all), nullptr, c.vn_nodes, c.n_vn);
    }
}

using CoopKernel = void (*)(IbArgs, CoopArgs);
CoopKernel coop_kernel_for(int dc, int dv, bool early, int T, bool match, int* smem_bytes);   // ib_n4_coop.cu; nullptr = not instantiated


// ------------------------------------------------------------------------------------------
// Irregular codes: the same whole-decode cooperative kernel with several degree classes per phase.  The classes of
// one phase are independent, so a phase is "for every class: stage its tables, __syncthreads, run its node list",
// with one grid-wide barrier per phase as before.  Instantiated for the degree sets of the reference's irregular
// codes (802.11n: d_c {7,8}, d_v {2,3,4,11}; DVB-S2: d_c {6,7}, d_v {1,2,3,8}); any other code keeps per-phase
// launches.  Variable nodes use the plain 2-word kernels (one CTA of 512 threads per SM, up to 128 registers).
// ------------------------------------------------------------------------------------------
constexpr int kCoopMaxClasses = 4;
struct CoopClasses {
    int n_cn_cls, n_vn_cls;
    int cn_deg[kCoopMaxClasses], vn_deg[kCoopMaxClasses];
    int cn_cnt[kCoopMaxClasses], vn_cnt[kCoopMaxClasses];
    const int* cn_nodes[kCoopMaxClasses];
    const int* vn_nodes[kCoopMaxClasses];
    const uint8_t* cn8;
    const uint8_t* vn8;
    const uint8_t* mc8;
    const uint8_t* mv8;
    const uint8_t* cn_pair;
    int DCmax, DVmax;
};

template <int... Ds> struct DegreeSet {};

template <int D>
__host__ __device__ constexpr int coop_cn_bytes(int T, bool match)
{
    return (D >= 6 ? (int)kPairBytes : 0) + n4_table_bytes(n4_cn_words(D, false)) + stage_scratch_bytes(D - 2, T, match ? D : 0);
}
template <int D>
__host__ __device__ constexpr int coop_vn_bytes(int T, bool match)
{
    const int up = n4_table_bytes(n4_vn_words(D, false)) + stage_scratch_bytes(D - 1, T, match ? D : 0);
    const int ou = n4_table_bytes(n4_vn_words(D, true)) + stage_scratch_bytes(D, T, 0);
    return up > ou ? up : ou;
}
constexpr int coop_max(int a, int b) { return a > b ? a : b; }
template <int... DCs, int... DVs>
constexpr int coop_multi_smem_bytes(DegreeSet<DCs...>, DegreeSet<DVs...>, int T, bool match)
{
    int m = 0;
    ((m = coop_max(m, coop_cn_bytes<DCs>(T, match))), ...);
    ((m = coop_max(m, coop_vn_bytes<DVs>(T, match))), ...);
    return m;
}

template <int D, bool EARLY>
__device__ __forceinline__ void coop_cn_class(const IbArgs& a, const CoopClasses& c, int ci, int it, uint32_t* s_all)
{
    constexpr bool PAIR = D >= 6;
    constexpr int NT = kCoopThreads;
    const int T = a.T, TT = T * T;
    uint32_t* s_cn = s_all + (PAIR ? kPairBytes / 4 : 0);
    IbArgs b = a;
    const int blk = it + 1;
    b.it = it; b.iter0 = (it < 0);
    b.lut = c.cn8 + (size_t)blk * (c.DCmax - 2) * TT;
    b.match = c.mc8 ? c.mc8 + (size_t)blk * c.DCmax * T : nullptr;
    b.nst = D - 2; b.dmax_match = c.mc8 ? D : 0;
    b.xp_col = PAIR ? D - 5 : -1;
    if (PAIR) {
        const uint2* src = reinterpret_cast<const uint2*>(c.cn_pair + ((size_t)blk * c.n_cn_cls + ci) * (size_t)TT * 8);
        uint2* dst = reinterpret_cast<uint2*>(s_all);
        for (int i = threadIdx.x; i < kTS * kTS * kPairSlots; i += NT) {
            const int r = i / kPairSlots, ra = r / kTS, rb = r - ra * kTS;
            dst[i] = (ra < T && rb < T) ? src[ra * T + rb] : make_uint2(0u, 0u);
        }
    }
    stage_tables_n4<n4_cn_words(D, false), NT>(s_cn, b, b.lut);
    __syncthreads();
    const uint32_t syn = a.B <= kLaneModeMaxFrames
                             ? cn_lanes_n4<D, false, EARLY, PAIR, NT>(b, reinterpret_cast<const uint8_t*>(s_cn),
                                                                      reinterpret_cast<const uint8_t*>(s_all), c.cn_nodes[ci], c.cn_cnt[ci])
                             : cn_loop_n4<D, false, EARLY, 2, PAIR, NT>(b, reinterpret_cast<const uint8_t*>(s_cn),
                                                                        reinterpret_cast<const uint8_t*>(s_all), c.cn_nodes[ci], c.cn_cnt[ci]);
    if (EARLY && it >= 0) {
        const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
        if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[it], 1);
    }
    __syncthreads();   // the next class restages the same shared memory
}

template <int D, bool DECIDE>
__device__ __forceinline__ void coop_vn_class(const IbArgs& a, const CoopClasses& c, int ci, int it, uint32_t* s_all)
{
    constexpr int NT = kCoopThreads;
    const int T = a.T, TT = T * T;
    IbArgs b = a;
    b.it = it; b.iter0 = 0;
    b.lut = c.vn8 + (size_t)it * c.DVmax * TT;
    b.match = (!DECIDE && c.mv8) ? c.mv8 + (size_t)it * c.DVmax * T : nullptr;
    b.nst = DECIDE ? D : D - 1; b.dmax_match = b.match ? D : 0;
    b.xp_col = -1;
    if (DECIDE || D > 1) {
        stage_tables_n4<n4_vn_words(D, DECIDE), NT>(s_all, b, b.lut);
        __syncthreads();
    }
    if (a.B <= kLaneModeMaxFrames) vn_lanes_n4<D, DECIDE, NT>(b, reinterpret_cast<const uint8_t*>(s_all), c.vn_nodes[ci], c.vn_cnt[ci]);
    else vn_loop_n4<D, DECIDE, 2, false, NT>(b, reinterpret_cast<const uint8_t*>(s_all), nullptr, c.vn_nodes[ci], c.vn_cnt[ci]);
    __syncthreads();
}

template <bool EARLY, int... DCs, int... DVs>
__device__ __forceinline__ void coop_multi_body(const IbArgs& a, const CoopClasses& c, uint32_t* s_all, DegreeSet<DCs...>,
                                                DegreeSet<DVs...>)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    auto cn_phase = [&](int it) {
        for (int ci = 0; ci < c.n_cn_cls; ++ci) {
            const int d = c.cn_deg[ci];
            ((d == DCs ? (coop_cn_class<DCs, EARLY>(a, c, ci, it, s_all), 0) : 0), ...);
        }
    };
    auto vn_phase = [&](int it) {
        for (int ci = 0; ci < c.n_vn_cls; ++ci) {
            const int d = c.vn_deg[ci];
            ((d == DVs ? (coop_vn_class<DVs, false>(a, c, ci, it, s_all), 0) : 0), ...);
        }
    };
    cn_phase(-1);
    grid.sync();
    int passes = 0;
    for (int it = 0; it < a.imax - 1; ++it) {
        if (EARLY && it >= 1 && *reinterpret_cast<volatile int*>(&a.flags[it - 1]) == 0) break;
        vn_phase(it);
        grid.sync();
        cn_phase(it);
        grid.sync();
        passes = it + 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.inum = passes + 1;
    for (int ci = 0; ci < c.n_vn_cls; ++ci) {
        const int d = c.vn_deg[ci];
        ((d == DVs ? (coop_vn_class<DVs, true>(a, c, ci, passes, s_all), 0) : 0), ...);
    }
}

template <typename CnSet, typename VnSet, bool EARLY>
__global__ void __launch_bounds__(kCoopThreads, 1) ib_decode_coop_multi_kernel(IbArgs a, CoopClasses c)
{
    extern __shared__ __align__(16) uint32_t s_all[];
    coop_multi_body<EARLY>(a, c, s_all, CnSet{}, VnSet{});
}

using CoopMultiKernel = void (*)(IbArgs, CoopClasses);
// ib_n4_coop.cu; nullptr = this combination of degree sets is not instantiated
CoopMultiKernel coop_multi_kernel_for(const int* cn_deg, int n_cn, const int* vn_deg, int n_vn, bool early, int T, bool match,
                                      int* smem_bytes);

}  // namespace ibldpc
