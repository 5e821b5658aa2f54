// ib_phase_n4.cuh -- fused per-phase kernels of the packed-nibble family: ONE launch per check-node / variable-node
// phase of the flooding schedule covering ALL degree classes of an irregular code.
//
// The per-class kernels of ib_kernels_n4.cuh pay a fixed cost per launch (launch latency, table staging through a
// scratch copy + striping loop + two block barriers, one partial wave at the end); an 802.11n decode is 301 launches
// of 60-160 us, and the small classes (d_c = 8: 108 checks, d_v = 4: 54 variables) cannot fill the tail of the big
// ones.  Here
//   * every CTA is identical: 1024 threads (32 warps at <= 64 registers), one CTA per SM, persistent;
//   * the shared-memory image of the phase -- lane-striped stage columns of every class side by side plus the
//     tail-pair rows -- is expanded ONCE on the host at ibldpc_set_luts (build_phase_images, ib_phase.cu) and
//     brought in by one thread with cp.async.bulk (TMA bulk copy, mbarrier complete_tx) while the warps are already
//     fetching the indices and messages of their first node;
//   * work = (class, node, frame tile) items, heaviest class first; a CTA owns a contiguous 1/gridDim share of every
//     class and its warps pull items from a shared-memory counter, so nothing waits for a slow class.
// The phase bodies are the per-class device functions (cn_word_n4_pair, vn_word_n4, vn_word_n4_pair) with the class's
// column base inside the shared image as a template parameter: same look-up order, bit-identical results
// (kernels_template_irreg.cl:33-99, :103-179, :181-246, :249-325).
//
// Instantiated per degree SET (DegreeSet<...>, heaviest first): the column layout is a compile-time function of the
// set, exactly like the multi-class cooperative kernel.  Other degree sets keep the per-class launches.
#pragma once
#include <cooperative_groups.h>

#include <utility>

#include "ib_coop_n4.cuh"   // DegreeSet
#include "ib_kernels_n4.cuh"
#include "ib_triple_n4.cuh"

namespace ibldpc {

constexpr int kPhaseThreads = 1024;
constexpr int kPhaseMaxClasses = 4;
constexpr int kPhaseCnPairMin = 6;   // tail-pair rows from this degree on (same thresholds as the per-class kernels)
constexpr int kPhaseVnPairMin = 5;

enum PhaseMode { kPhaseCn = 0, kPhaseVn = 1, kPhaseOut = 2 };

// ---- compile-time layout of a phase image -------------------------------------------------------------------
// A class of a degree set is written as its degree d (default composition: tail pair from degree 6 / 5 on) or as
// d + 100 * (pm + 1) with an explicit pair mode pm: 0 = plain chains, 1 = tail pair.  The composed table costs 32 KB of
// the phase image, which a class of a handful of nodes (DVB-S2: one check of degree 6) is not worth.
// (A second composed row in the MIDDLE of the chains, H(m_{D-4}, m_{D-3})[t], was implemented and measured in round 2: it
// removes 17-24 % of the shared-memory wavefronts, is bit-exact, and is SLOWER -- (3,6) check-node phase 0.48 -> 0.56 ms,
// DVB-S2 0.67 -> 0.72 ms, 802.11n 0.178 -> 0.189 ms: the 64-bit nibble selects cost more issue slots than the look-ups
// they replace, and issue is the co-limiter.  profiles/README.md.)
__host__ __device__ constexpr int spec_deg(int v) { return v % 100; }
__host__ __device__ constexpr int phase_cols(int mode, int d) { return mode == kPhaseCn ? d - 2 : mode == kPhaseVn ? d - 1 : d; }
__host__ __device__ constexpr int phase_pm(int mode, int v)
{
    if (mode == kPhaseOut) return 0;
    if (v >= 100) return v / 100 - 1;
    return mode == kPhaseCn ? (v >= kPhaseCnPairMin ? 1 : 0) : (v >= kPhaseVnPairMin ? 1 : 0);
}
// pm = 2 (written d + 300): the three-input table of ib_triple_n4.cuh -- check nodes of degree 6 (next to the tail-pair rows),
// variable nodes of degree 3 (instead of any stage column); single-class sets only (the table is 128 KB).
__host__ __device__ constexpr bool phase_pair(int mode, int v) { return mode == kPhaseCn ? phase_pm(mode, v) >= 1 : phase_pm(mode, v) == 1; }
__host__ __device__ constexpr bool phase_tri(int mode, int v) { return phase_pm(mode, v) == 2; }
// words (8 frames each) a lane moves per message row: the widest access the register budget of 64 allows
__host__ __device__ constexpr int phase_vec(int mode, int d)
{
    return mode == kPhaseCn ? 2 : mode == kPhaseVn ? (d <= 4 ? 4 : d <= 7 ? 2 : 1) : 2;
}

template <int MODE, int... Ds>
struct PhaseLayout {
    static constexpr int n = sizeof...(Ds);
    __host__ __device__ static constexpr int spec(int i)
    {
        constexpr int d[] = {Ds...};
        return d[i];
    }
    __host__ __device__ static constexpr int degree(int i) { return spec_deg(spec(i)); }
    __host__ __device__ static constexpr int pm(int i) { return phase_pm(MODE, spec(i)); }
    // first stage column of class i (the decision phase shares its plain columns between the classes)
    __host__ __device__ static constexpr int col_base(int i)
    {
        if (MODE == kPhaseOut) return 0;
        int c = 0;
        for (int j = 0; j < i; ++j) c += phase_cols(MODE, degree(j));
        return c;
    }
    __host__ __device__ static constexpr int total_cols()
    {
        int c = 0;
        for (int j = 0; j < n; ++j) c = MODE == kPhaseOut ? (degree(j) > c ? degree(j) : c) : c + phase_cols(MODE, degree(j));
        return c;
    }
    static constexpr int words = total_cols() <= 4 ? 1 : (total_cols() + 3) / 4;
    __host__ __device__ static constexpr bool pair(int i) { return phase_pair(MODE, spec(i)); }
    __host__ __device__ static constexpr bool tri(int i) { return phase_tri(MODE, spec(i)); }
    __host__ __device__ static constexpr int pair_index(int i)
    {
        int p = 0;
        for (int j = 0; j < i; ++j) p += pair(j) ? 1 : 0;
        return p;
    }
    __host__ __device__ static constexpr int tri_count()
    {
        int p = 0;
        for (int j = 0; j < n; ++j) p += tri(j) ? 1 : 0;
        return p;
    }
    static constexpr int n_pair = pair_index(n);
    static constexpr int n_tri = tri_count();
    static_assert(n_tri == 0 || n == 1, "three-input table: single-class degree sets only");
    // image: [tail-pair rows of the classes][three-input table][stage columns]
    static constexpr int tri_offset = n_pair * (int)kPairBytes;
    static constexpr int tab_offset = tri_offset + n_tri * kTripleBytes;
    static constexpr int image_bytes = tab_offset + n4_table_bytes(words);
};

// run-time description of the same layout for the host-side image builder (ib_phase.cu)
struct PhaseClassLayout { int degree, col_base, cols, pair_index, vec, pm; bool pair, tri; };
struct PhaseLayoutRt { int n, words, n_pair, image_bytes, tri_offset, tab_offset; PhaseClassLayout cls[kPhaseMaxClasses]; };

template <int MODE, int... Ds>
PhaseLayoutRt phase_layout_rt(DegreeSet<Ds...>)
{
    using L = PhaseLayout<MODE, Ds...>;
    PhaseLayoutRt r{};
    r.n = L::n; r.words = L::words; r.n_pair = L::n_pair; r.image_bytes = L::image_bytes;
    r.tri_offset = L::tri_offset; r.tab_offset = L::tab_offset;
    for (int i = 0; i < L::n; ++i) {
        const int d = L::degree(i);
        r.cls[i] = PhaseClassLayout{d, L::col_base(i), phase_cols(MODE, d), L::pair_index(i), phase_vec(MODE, d), L::pm(i), L::pair(i), L::tri(i)};
    }
    return r;
}

struct PhaseArgs {
    IbArgs a;                    // graph, buffers, iteration control (lut / match / pair fields unused)
    const uint8_t* image;        // shared-memory image of this phase in global memory (16-byte aligned)
    long long image_stride;      // decision phase: bytes between the images of consecutive iterations
    const int* nodes[kPhaseMaxClasses];    // node ids of every class, in the order of the degree set
    const int* starts[kPhaseMaxClasses];   // sc[node] resp. sv[node] of the same nodes (saves one dependent load)
    int n_nodes[kPhaseMaxClasses];
    // per-frame early termination with frame compaction (ib_perframe.cu); unused (null) otherwise
    const struct PfState* pf;       // device-side state: active columns, done flag
    const uint32_t* pf_alive;       // [words] nibble mask of the frames that still iterate (the others are frozen)
    uint32_t* pf_fsyn;              // [words] bit 4f: frame f of the word failed a check in this pass
    // deferred decision of the converged frames (ib_phase_pfdecide_kernel)
    const int* pf_fin;              // [slots] column of the current buffers behind every dense result slot (-1 = padding)
    const int* pf_gstart;           // [imax + 1] first slot of the frames that converged in pass g - 1 (group g)
    uint8_t* pf_res;                // [n_var][pf_res_pitch] decided cluster indices as nibbles, one per result slot
    uint32_t pf_res_pitch;
};

// Device-side state of a per-frame-early-termination decode: every kernel of the schedule is launched unconditionally
// and reads what is left to do from here (no host synchronisation inside a decode).
struct PfState {
    int n_act;        // columns [0, n_act) of the buffers are in use (alive, or converged since the last compaction)
    int act_pitch;    // ceil(n_act / 2) rounded up to 16 bytes
    int n_holes;      // columns the next compaction moves (dead columns of the new front = alive columns behind it)
    int n_alive;      // frames still iterating
    int done;         // n_alive == 0: every later phase kernel returns at once
    int do_compact;   // set by pf_scan_kernel when the next gather is worth it
    int new_n;        // alive columns after that gather
    int alive_acc;    // accumulator of pf_update_kernel
    unsigned blocks_done;
    unsigned blocks_done2;   // pf_commit_kernel
    int fin_count;    // dense result slots handed out so far
    float waste;      // passes' worth of work spent on converged columns since the last compaction
};

// ---- TMA bulk copy of the image ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void phase_image_issue(void* smem_dst, const uint8_t* gsrc, uint32_t bytes, uint64_t* mbar)
{
    const uint32_t mb = smem_u32(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
    constexpr uint32_t kChunk = 32768;   // several bulk operations in flight
    const uint32_t dst = smem_u32(smem_dst);
    for (uint32_t off = 0; off < bytes; off += kChunk) {
        const uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                     "l"(gsrc + off), "r"(n), "r"(mb)
                     : "memory");
    }
}

__device__ __forceinline__ void phase_image_wait(uint64_t* mbar)
{
    const uint32_t mb = smem_u32(mbar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mb)
            : "memory");
    }
}

// a warp-uniform int re-read through the read-only path WITHOUT common-subexpression elimination against an earlier
// load of the same address: lets the high-degree classes drop their row indices during the look-up chains and fetch
// them again (L1 hits) for the stores, instead of holding D registers or spilling
__device__ __forceinline__ int ld_nc_again(const int* p)
{
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// ---- one item = one (node, tile) of class I ------------------------------------------------------------------
struct PfCtx { uint32_t* fsyn; const uint32_t* alive; uint32_t fsyn_s; };   // fsyn_s: shared address of the per-CTA accumulator

template <int MODE, bool EARLY, typename L, int I, int PF = 0>   // PF: per-frame early termination, 1 = syndrome flags in shared memory, 2 = in global memory
struct PhaseItem {
    static constexpr int D = L::degree(I);
    static constexpr int VEC = phase_vec(MODE, D);
    static constexpr int WT = L::words;
    static constexpr int CB = L::col_base(I);
    static constexpr int PM = L::pm(I);
    static constexpr bool PAIR = L::pair(I);
    static constexpr bool TRI = L::tri(I);
    static_assert(!TRI || (MODE == kPhaseCn && D == 6 && L::words == 1 && L::col_base(I) == 0) || (MODE == kPhaseVn && D == 3),
                  "three-input table: check nodes of degree 6 / variable nodes of degree 3");
    static constexpr int PI = L::pair_index(I);

    // returns the syndrome bits seen (check-node phase with EARLY), 0 otherwise
    static __device__ __forceinline__ uint32_t run(const IbArgs& a, const uint8_t* s_img, int node, int start, uint32_t col,
                                                   uint32_t lane4, const PfCtx& pf)
    {
        const uint8_t* tab = s_img + L::tab_offset;
        const uint8_t* ptab = s_img + PI * kPairBytes;   // the three-input table of a check-node class follows its pair rows
        if constexpr (MODE == kPhaseCn) {
            return cn_node_n4<D, false, EARLY, VEC, PAIR, WT, CB, PF, TRI ? 1 : 0>(a, tab, ptab, start, col, lane4, a.B - 2 * (int)col,
                                                                      PF ? pf.fsyn + (col >> 2) : nullptr,
                                                                      PF ? pf.alive + (col >> 2) : nullptr, PF ? pf.fsyn_s + col : 0u);
        } else {
            constexpr bool DECIDE = MODE == kPhaseOut;
            constexpr bool kKeepRows = D <= 6;   // row indices stay in registers only where the budget allows
            int rows[D];
            VnIn4<D, VEC> in;
            ld_words<VEC>(a.ch + (uint64_t)(uint32_t)node * a.pitch + col, in.c);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                rows[k] = a.tv[start + k];
                ld_words<VEC>(a.msg + (uint64_t)(uint32_t)rows[k] * a.pitch + col, in.m[k]);
            }
            if (!DECIDE && D == 1) {   // degree-1 variable node forwards the raw channel value (:132-136)
                st_words<VEC>(a.msg + (uint64_t)(uint32_t)rows[0] * a.pitch + col, in.c);
                return 0;
            }
            uint32_t r[D][VEC];
            uint32_t dec[2 * VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                uint32_t w[D], o[D];
#pragma unroll
                for (int k = 0; k < D; ++k) w[k] = in.m[k][j];
                if constexpr (TRI && D == 3) vn3_word_n4(in.c[j], w, o, s_img + L::tri_offset, lane4);
                else if constexpr (PAIR) vn_word_n4_pair<D, WT, CB>(in.c[j], w, o, tab, ptab, lane4, (lane4 & (4u * (kPairSlots - 1))) * 2u);
                else vn_word_n4<D, DECIDE, WT, CB>(in.c[j], w, o, dec[2 * j], dec[2 * j + 1], tab, lane4);
                if (!DECIDE) {
                    if constexpr (PF) {
                        // converged frames keep the check-to-variable messages they converged with (decided later)
                        const uint32_t am = pf.alive[(col >> 2) + j];
#pragma unroll
                        for (int k = 0; k < D; ++k) o[k] = (o[k] & am) | (w[k] & ~am);
                    }
#pragma unroll
                    for (int k = 0; k < D; ++k) r[k][j] = o[k];
                }
            }
            if constexpr (DECIDE) {
                // decided cluster indices leave as uint8 (one byte per frame): 8*VEC bytes per lane
                const uint32_t ocol = 2u * col;
                uint8_t* dst = a.out + (uint64_t)(uint32_t)node * a.out_pitch + ocol;
                if constexpr (VEC >= 2) {
#pragma unroll
                    for (int q = 0; q < VEC / 2; ++q)
                        if (ocol + 16u * q < a.out_pitch)
                            *reinterpret_cast<uint4*>(dst + 16 * q) = make_uint4(dec[4 * q], dec[4 * q + 1], dec[4 * q + 2], dec[4 * q + 3]);
                } else {
                    if (ocol < a.out_pitch) *reinterpret_cast<uint2*>(dst) = make_uint2(dec[0], dec[1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const int row = kKeepRows ? rows[k] : ld_nc_again(a.tv + start + k);
                    st_words<VEC>(a.msg + (uint64_t)(uint32_t)row * a.pitch + col, r[k]);
                }
            }
            return 0;
        }
    }
};

template <typename F, int... Is>
__device__ __forceinline__ void phase_unroll(F& f, std::integer_sequence<int, Is...>)
{
    (f(std::integral_constant<int, Is>{}), ...);
}

// ---- the kernel -----------------------------------------------------------------------------------------------
template <int MODE, bool EARLY, int PF, int... Ds>
__device__ __forceinline__ void phase_kernel_body(const PhaseArgs& p, const IbArgs& a, uint32_t bound, const PfCtx& pfc)
{
    using L = PhaseLayout<MODE, Ds...>;
    static_assert(L::n <= kPhaseMaxClasses, "too many degree classes");
    extern __shared__ __align__(128) uint8_t s_img[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ int s_next[kPhaseMaxClasses];
    __shared__ int s_passes;

    const int lane = threadIdx.x & 31;
    const uint32_t lane4 = lane * 4;
    // this CTA's contiguous share [lo, hi) of every class (items = nodes x tiles)
    int lo[L::n], hi[L::n], tiles[L::n];
    {
        constexpr int degs[] = {Ds...};
#pragma unroll
        for (int c = 0; c < L::n; ++c) {
            const int vec = phase_vec(MODE, spec_deg(degs[c]));
            tiles[c] = (int)((bound + 128u * vec - 1) / (128u * vec));
            const long long items = (long long)p.n_nodes[c] * tiles[c];
            lo[c] = (int)(items * blockIdx.x / gridDim.x);
            hi[c] = (int)(items * (blockIdx.x + 1) / gridDim.x);
        }
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < L::n; ++c) s_next[c] = lo[c] + kPhaseThreads / 32;   // the first 32 items are taken statically
        const uint8_t* img = p.image;
        if (MODE == kPhaseOut) {
            s_passes = executed_passes(a);
            if (blockIdx.x == 0) *a.inum = s_passes + 1;
            img += (long long)s_passes * p.image_stride;
        }
        phase_image_issue(s_img, img, (uint32_t)L::image_bytes, &s_mbar);
    }
    __syncthreads();   // mbarrier initialised, counters set; the image is still in flight

    uint32_t syn = 0;
    bool have_image = false;
    constexpr int degs[] = {Ds...};
    auto class_loop = [&](auto IC) {
        constexpr int I = decltype(IC)::value;
        using Item = PhaseItem<MODE, EARLY, L, I, PF>;
        constexpr int VEC = Item::VEC;
        const int* __restrict__ nodes = p.nodes[I];
        const int* __restrict__ starts = p.starts[I];
        const uint32_t tl = (uint32_t)tiles[I];
        int i = lo[I] + (threadIdx.x >> 5);
        int node = 0, start = 0;
        if (i < hi[I]) {
            const uint32_t ni = (uint32_t)i / tl;
            node = nodes[ni];
            start = starts[ni];
        }
        while (i < hi[I]) {
            // take the next item and fetch its node before working on this one
            int i2 = 0;
            if (lane == 0) i2 = atomicAdd(&s_next[I], 1);
            i2 = __shfl_sync(0xffffffffu, i2, 0);
            int node2 = 0, start2 = 0;
            if (i2 < hi[I]) {
                const uint32_t ni2 = (uint32_t)i2 / tl;
                node2 = nodes[ni2];
                start2 = starts[ni2];
            }
            const uint32_t tile = (uint32_t)i - ((uint32_t)i / tl) * tl;
            const uint32_t col = (tile * 32u + lane) * (4u * VEC);
            if (!have_image) {        // first item of this warp: its row indices are on their way, now wait for the tables
                phase_image_wait(&s_mbar);
                have_image = true;
            }
            if (col < bound) syn |= Item::run(a, s_img, node, start, col, lane4, pfc);
            i = i2;
            node = node2;
            start = start2;
        }
        (void)degs;
    };
    // classes in the order of the degree set (heaviest first)
    phase_unroll(class_loop, std::make_integer_sequence<int, L::n>{});

    if (MODE == kPhaseCn && EARLY && !PF && !a.iter0) {
        const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
        if (any != 0 && lane == 0) atomicOr(&a.flags[a.it], 1);
    }
    // the bulk copy must have landed before the CTA may exit (its shared memory is the destination)
    if (!have_image) phase_image_wait(&s_mbar);
}

template <int MODE, bool EARLY, int... Ds>
__global__ void __launch_bounds__(kPhaseThreads, 1) ib_phase_kernel(PhaseArgs p)
{
    const IbArgs& a = p.a;
    if (MODE != kPhaseOut && (EARLY || a.early) && a.it >= 1 && a.flags[a.it - 1] == 0) return;   // batch already converged
    phase_kernel_body<MODE, EARLY, 0, Ds...>(p, a, a.pitch, PfCtx{nullptr, nullptr, 0u});
}

// Per-frame early termination: same bodies over the ACTIVE columns; the check-node phase records which frames failed a
// check (pf_fsyn), and both phases leave the messages of the frames that have converged untouched (pf_alive).
// SYN (check-node phase): 1 = per-CTA syndrome accumulator in shared memory behind the table image (one global atomicOr
// per word and CTA at the end instead of one per check and word; the host checks that the active words fit), 2 = global.
template <int MODE, int SYN, int... Ds>
__global__ void __launch_bounds__(kPhaseThreads, 1) ib_phase_pf_kernel(PhaseArgs p)
{
    static_assert(MODE != kPhaseOut, "converged frames are decided by ib_phase_pfdecide_kernel");
    using L = PhaseLayout<MODE, Ds...>;
    extern __shared__ __align__(128) uint8_t s_img[];
    const int done = p.pf->done, n_act = p.pf->n_act, act_pitch = p.pf->act_pitch;
    if (done || n_act == 0) return;
    IbArgs a = p.a;
    a.B = n_act;
    const int act_words = act_pitch >> 2;
    uint32_t* s_fs = reinterpret_cast<uint32_t*>(s_img + L::image_bytes);
    if (MODE == kPhaseCn && SYN == 1)
        for (int w = threadIdx.x; w < act_words; w += kPhaseThreads) s_fs[w] = 0u;   // ordered by the body's first barrier
    phase_kernel_body<MODE, MODE == kPhaseCn, SYN, Ds...>(p, a, (uint32_t)act_pitch,
                                                         PfCtx{p.pf_fsyn, p.pf_alive, smem_u32(s_img) + (uint32_t)L::image_bytes});
    if (MODE == kPhaseCn && SYN == 1) {
        __syncthreads();
        for (int w = threadIdx.x; w < act_words; w += kPhaseThreads) {
            const uint32_t v = s_fs[w];
            if (v) atomicOr(p.pf_fsyn + w, v);
        }
    }
}

// ---- deferred decision of the frames that converged (per-frame early termination) ----------------------------
// Group g = the frames whose syndrome became zero in pass g - 1 (group 0: no pass at all, i_max <= 1); it owns the dense
// result slots [gstart[g], gstart[g + 1]) (a multiple of 8), and pf_fin names, for every slot, the column of the message
// and channel arrays in which that frame sits FROZEN since it converged.  One launch decides all groups [g_lo, g_hi]: for
// every group the decision image of iteration g (calc_varnode_output with the tables of i_num - 1,
// discrete_LDPC_decoder.py:280-287) is brought in by TMA, and one lane gathers the eight frames of a result word from
// their columns, runs the same decision chain as ib_phase_kernel<kPhaseOut> and stores the word: every frame is decided
// exactly once, densely, whatever the order in which the frames of a batch converge.
template <int D>
__device__ __forceinline__ uint32_t pf_gather8(const uint8_t* row, const int (&cols)[8])
{
    const uint32_t* r32 = reinterpret_cast<const uint32_t*>(row);
    uint32_t v = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) v |= ((r32[cols[q] >> 3] >> (4 * (cols[q] & 7))) & 15u) << (4 * q);
    return v;
}

template <typename L, int I>
struct PfDecideItem {
    static constexpr int D = L::degree(I);
    static __device__ __forceinline__ uint32_t run(const IbArgs& a, const uint8_t* s_img, int node, int start, const int (&cols)[8],
                                                   uint32_t lane4)
    {
        const uint8_t* tab = s_img + L::tab_offset;
        const uint32_t chw = pf_gather8<D>(a.ch + (uint64_t)(uint32_t)node * a.pitch, cols);
        uint32_t w[D], o[D], dlo, dhi;
#pragma unroll
        for (int k = 0; k < D; ++k) w[k] = pf_gather8<D>(a.msg + (uint64_t)(uint32_t)a.tv[start + k] * a.pitch, cols);
        vn_word_n4<D, true, L::words, L::col_base(I)>(chw, w, o, dlo, dhi, tab, lane4);
        const uint32_t lo = dlo & 0x0f0f0f0fu, hi = dhi & 0x0f0f0f0fu;
        const uint32_t l2 = (lo | (lo >> 4)) & 0x00ff00ffu, h2 = (hi | (hi >> 4)) & 0x00ff00ffu;
        return ((l2 | (l2 >> 8)) & 0xffffu) | (((h2 | (h2 >> 8)) & 0xffffu) << 16);
    }
};

template <int... Ds>
__global__ void __launch_bounds__(kPhaseThreads, 1) ib_phase_pfdecide_kernel(PhaseArgs p, int g_lo, int g_hi)
{
    using L = PhaseLayout<kPhaseOut, Ds...>;
    extern __shared__ __align__(128) uint8_t s_img[];
    __shared__ __align__(8) uint64_t s_mbar;
    const int* __restrict__ gstart = p.pf_gstart;
    if (gstart[g_hi + 1] == gstart[g_lo]) return;   // nothing converged in these passes
    const IbArgs& a = p.a;
    const int lane = threadIdx.x & 31;
    const uint32_t lane4 = lane * 4;
    const int gw = blockIdx.x * (kPhaseThreads / 32) + (threadIdx.x >> 5), n_gw = gridDim.x * (kPhaseThreads / 32);
    const uint32_t mb = smem_u32(&s_mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t parity = 0;
    for (int g = g_lo; g <= g_hi; ++g) {
        const int s0 = gstart[g], s1 = gstart[g + 1];
        if (s1 == s0) continue;
        __syncthreads();   // mbarrier initialised / nobody reads the previous image any more
        if (threadIdx.x == 0) {
            const uint8_t* img = p.image + (long long)g * p.image_stride;
            constexpr uint32_t bytes = (uint32_t)L::image_bytes, kChunk = 32768;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            const uint32_t dst = smem_u32(s_img);
            for (uint32_t off = 0; off < bytes; off += kChunk) {
                const uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                             "l"(img + off), "r"(n), "r"(mb)
                             : "memory");
            }
        }
        const int words_g = (s1 - s0) >> 3, tiles = (words_g + 31) >> 5;
        {   // wait for the image (phase parity alternates per group)
            uint32_t ok = 0;
            while (!ok) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(mb), "r"(parity)
                    : "memory");
            }
            parity ^= 1u;
        }
        auto class_loop = [&](auto IC) {
            constexpr int I = decltype(IC)::value;
            const int* __restrict__ nodes = p.nodes[I];
            const int* __restrict__ starts = p.starts[I];
            const long long items = (long long)p.n_nodes[I] * tiles;
            for (long long i = gw; i < items; i += n_gw) {
                const int ni = (int)(i / tiles), wd = (int)(i - (long long)ni * tiles) * 32 + lane;
                if (wd >= words_g) continue;
                const int4* f4 = reinterpret_cast<const int4*>(p.pf_fin + s0 + 8 * wd);
                const int4 fa = f4[0], fb = f4[1];
                const int cols[8] = {max(fa.x, 0), max(fa.y, 0), max(fa.z, 0), max(fa.w, 0),
                                     max(fb.x, 0), max(fb.y, 0), max(fb.z, 0), max(fb.w, 0)};   // padding slots read column 0
                const int node = nodes[ni];
                const uint32_t val = PfDecideItem<L, I>::run(a, s_img, node, starts[ni], cols, lane4);
                reinterpret_cast<uint32_t*>(p.pf_res + (uint64_t)(uint32_t)node * p.pf_res_pitch)[(s0 >> 3) + wd] = val;
            }
        };
        phase_unroll(class_loop, std::make_integer_sequence<int, L::n>{});
    }
}


// ---- small batches: the whole decode in ONE cooperative launch over the phase images ---------------------------
// B <= kLaneModeMaxFrames (256): a lane is a (node, word) pair (cn_lanes_n4 / vn_lanes_n4), the whole flooding schedule
// runs in one cooperative launch with a grid-wide barrier between the phases (ib_coop_n4.cuh) -- but where those kernels
// restage the tables of every degree class of every phase with a striping loop and two block barriers (six times per
// iteration for the DVB-S2 set: 39 us per iteration at B = 2, the reference's msg_at_time), this one brings in the
// pre-expanded image of the WHOLE phase with one TMA bulk copy issued BEFORE the grid barrier, so the copy overlaps the
// barrier and all classes of the phase run back to back without a block barrier in between.
// Same look-up functions with the class's column base inside the image: bit-identical results, same stop rule, same i_num.
struct CoopPhaseArgs {
    IbArgs a;
    const uint8_t* cn_images;   // [imax] check-node images (block 0 = iteration-0 tables)
    const uint8_t* vn_images;   // [imax] variable-node update images
    const uint8_t* out_images;  // [imax] decision images
    const int* cn_nodes[kPhaseMaxClasses];
    const int* vn_nodes[kPhaseMaxClasses];
    const int* cn_starts[kPhaseMaxClasses];   // sc[node] / sv[node] of the same nodes (warp-per-(node, tile) bodies)
    const int* vn_starts[kPhaseMaxClasses];
    int cn_count[kPhaseMaxClasses], vn_count[kPhaseMaxClasses];
};

// one class of one phase with a warp per (node, tile) item, items dealt round-robin to the warps of the grid
// (batches above the lane mode: kLaneModeMaxFrames < B <= coop_max_frames)
template <int NT, typename Item>
__device__ __forceinline__ uint32_t coop_phase_tiles(const IbArgs& b, const uint8_t* s_img, const int* __restrict__ nodes,
                                                     const int* __restrict__ starts, int n_nodes)
{
    constexpr int VEC = Item::VEC;
    const int lane = threadIdx.x & 31;
    const uint32_t tiles = (uint32_t)((b.pitch + 128u * VEC - 1) / (128u * VEC));
    const long long items = (long long)n_nodes * tiles;
    const long long gw = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (NT / 32);
    uint32_t syn = 0;
    for (long long i = gw; i < items; i += nw) {
        const uint32_t ni = (uint32_t)(i / tiles), tile = (uint32_t)(i - (long long)ni * tiles);
        const uint32_t col = (tile * 32u + lane) * (4u * VEC);
        if (col < b.pitch) syn |= Item::run(b, s_img, nodes[ni], starts[ni], col, lane * 4u, PfCtx{nullptr, nullptr, 0u});
    }
    return syn;
}

template <int NT, bool EARLY, int... Cs, int... Vs>
__device__ __forceinline__ void coop_phase_body(const CoopPhaseArgs& p, DegreeSet<Cs...>, DegreeSet<Vs...>)
{
    namespace cg = cooperative_groups;
    using LC = PhaseLayout<kPhaseCn, Cs...>;
    using LV = PhaseLayout<kPhaseVn, Vs...>;
    using LO = PhaseLayout<kPhaseOut, Vs...>;
    extern __shared__ __align__(128) uint8_t s_img[];
    __shared__ __align__(8) uint64_t s_mbar;
    cg::grid_group grid = cg::this_grid();
    const IbArgs& a = p.a;
    const uint32_t mb = smem_u32(&s_mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t parity = 0;
    const bool lanes = a.B <= kLaneModeMaxFrames;   // lane = (node, word); above: warp = (node, tile)
    // every thread of the CTA has left the previous image (and seen its mbarrier phase) before the next copy is issued
    auto issue = [&](const uint8_t* img, uint32_t bytes) {
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the old image before the async-proxy writes
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            constexpr uint32_t kChunk = 32768;
            const uint32_t dst = smem_u32(s_img);
            for (uint32_t off = 0; off < bytes; off += kChunk) {
                const uint32_t n = bytes - off < kChunk ? bytes - off : kChunk;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + off),
                             "l"(img + off), "r"(n), "r"(mb)
                             : "memory");
            }
        }
    };
    auto wait = [&]() {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok)
                : "r"(mb), "r"(parity)
                : "memory");
        }
        parity ^= 1u;
    };
    auto cn_phase = [&](int it) {
        IbArgs b = a;
        b.it = it; b.iter0 = (it < 0);
        uint32_t syn = 0;
        auto one = [&](auto IC) {
            constexpr int I = decltype(IC)::value;
            constexpr int D = LC::degree(I);
            constexpr bool PAIR = LC::pair(I);
            if (lanes)
                syn |= cn_lanes_n4<D, false, EARLY, PAIR, NT, LC::words, LC::col_base(I)>(
                    b, s_img + LC::tab_offset, s_img + LC::pair_index(I) * kPairBytes, p.cn_nodes[I], p.cn_count[I]);
            else
                syn |= coop_phase_tiles<NT, PhaseItem<kPhaseCn, EARLY, LC, I, 0>>(b, s_img, p.cn_nodes[I], p.cn_starts[I], p.cn_count[I]);
        };
        phase_unroll(one, std::make_integer_sequence<int, LC::n>{});
        if (EARLY && it >= 0) {
            const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
            if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[it], 1);
        }
    };
    auto vn_phase = [&](int it) {
        IbArgs b = a;
        b.it = it; b.iter0 = 0;
        auto one = [&](auto IC) {
            constexpr int I = decltype(IC)::value;
            constexpr int D = LV::degree(I);
            constexpr bool PAIR = LV::pair(I);
            if (lanes)
                vn_lanes_n4<D, false, NT, LV::words, LV::col_base(I), PAIR>(b, s_img + LV::tab_offset, p.vn_nodes[I], p.vn_count[I],
                                                                            s_img + LV::pair_index(I) * kPairBytes);
            else
                coop_phase_tiles<NT, PhaseItem<kPhaseVn, false, LV, I, 0>>(b, s_img, p.vn_nodes[I], p.vn_starts[I], p.vn_count[I]);
        };
        phase_unroll(one, std::make_integer_sequence<int, LV::n>{});
    };
    auto out_phase = [&](int it) {
        IbArgs b = a;
        b.it = it; b.iter0 = 0;
        auto one = [&](auto IC) {
            constexpr int I = decltype(IC)::value;
            if (lanes)
                vn_lanes_n4<LO::degree(I), true, NT, LO::words, LO::col_base(I), false>(b, s_img + LO::tab_offset, p.vn_nodes[I],
                                                                                       p.vn_count[I]);
            else
                coop_phase_tiles<NT, PhaseItem<kPhaseOut, false, LO, I, 0>>(b, s_img, p.vn_nodes[I], p.vn_starts[I], p.vn_count[I]);
        };
        phase_unroll(one, std::make_integer_sequence<int, LO::n>{});
    };

    issue(p.cn_images, (uint32_t)LC::image_bytes);
    wait();
    cn_phase(-1);
    int passes = 0;
    bool converged = false;
    for (int it = 0; it < a.imax - 1; ++it) {
        issue(p.vn_images + (size_t)it * LV::image_bytes, (uint32_t)LV::image_bytes);   // overlaps the grid barrier
        grid.sync();
        wait();
        // reference stop rule (discrete_LDPC_decoder.py:233-276): pass `it` runs iff it == 0 or pass it-1 left a non-zero
        // syndrome somewhere in the batch; flags[] were written before the grid barrier
        if (EARLY && it >= 1 && *reinterpret_cast<volatile int*>(&a.flags[it - 1]) == 0) { converged = true; break; }
        vn_phase(it);
        issue(p.cn_images + (size_t)(it + 1) * LC::image_bytes, (uint32_t)LC::image_bytes);
        grid.sync();
        wait();
        cn_phase(it);
        passes = it + 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.inum = passes + 1;
    // calc_varnode_output with the tables of iteration i_num - 1; after a converged stop the grid is already in step
    issue(p.out_images + (size_t)passes * LO::image_bytes, (uint32_t)LO::image_bytes);
    if (!converged) grid.sync();
    wait();
    out_phase(passes);
}

template <int NT, bool EARLY, typename CnSet, typename VnSet>
__global__ void __launch_bounds__(NT, 1) ib_coop_phase_kernel(CoopPhaseArgs p)
{
    coop_phase_body<NT, EARLY>(p, CnSet{}, VnSet{});
}

using CoopPhaseKernel = void (*)(CoopPhaseArgs);

using PhaseKernel = void (*)(PhaseArgs);
using PhaseDecideKernel = void (*)(PhaseArgs, int, int);

}  // namespace ibldpc
