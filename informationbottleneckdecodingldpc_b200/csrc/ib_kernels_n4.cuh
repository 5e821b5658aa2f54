// ib_kernels_n4.cuh -- packed-nibble variant of the IB fast path (|T| <= 16, T even).
//
// Same flooding schedule, same in-place check-node-major message array and the same look-up
// order as ib_kernels.cuh (so results stay bit-identical to the reference chains,
// kernels_template_irreg.cl:33-99, :103-179, :181-246, :249-325), but every message and channel
// value is stored in FOUR bits:
//   ch4 [n_var ][pitch4]   channel cluster indices, two frames per byte
//   msg [n_edge][pitch4]   messages, two frames per byte
// frame F of a row sits in bits 4*(F&7).. of the 32-bit word F>>3, pitch4 = ceil(B/2) rounded up
// to 16 bytes.  HBM traffic per iteration drops from 4E+N to 2E+N/2 bytes per frame (SURVEY 8d
// names nibble packing as the lever above the uint8 roofline).
//
// Thread mapping: one warp = one (node, tile); a lane moves VEC 32-bit words (8*VEC frames) per
// message with one 64- or 128-bit access, so a warp touches 128*VEC contiguous bytes per row.
//
// Tables: lane-striped shared-memory layout of ib_kernels.cuh with a fixed row stride of 16
// values per message (row = m*16 + t), so every address term is a shift by a compile-time amount:
//   addr = (t << log2(128 W)) + (m << log2(2048 W)) + lane*4 + column offset.
#pragma once
#include "ib_kernels.cuh"

// Address arithmetic of the look-ups: IDP.4A on the fma pipe (default, see "dp4a address arithmetic" below) or shifts and
// LOP3 on the alu pipe (-DIBLDPC_NO_DP4A: the round-1 form, kept for A/B builds and for the one kernel family that is
// faster with it, ib_n4_vn_pair.cu).  The two forms use different table-row orders, so the switch is per translation
// unit: a kernel stages its own tables (stage_tables_n4) or reads images built in the same mode (ib_phase.cu).
#if !defined(IBLDPC_NO_DP4A) && !defined(IBLDPC_DP4A)
#define IBLDPC_DP4A 1
#endif

namespace ibldpc {

constexpr int kTS = 16;   // row stride per message value of the n4 table layout

__host__ __device__ constexpr int n4_words(int cols) { return cols <= 4 ? 1 : (cols + 3) / 4; }
__host__ __device__ constexpr int n4_cn_words(int d, bool explicit_match) { return n4_words(d - 2 + (explicit_match ? 1 : 0)); }
__host__ __device__ constexpr int n4_vn_words(int d, bool decide) { return n4_words(decide ? d : d - 1); }
__host__ __device__ constexpr int n4_table_bytes(int W) { return kTS * kTS * W * 128; }

// ------------------------------------------------------------------------------------------
// word-vector loads / stores (VEC = 2: 64-bit, VEC = 4: 128-bit)
// ------------------------------------------------------------------------------------------
template <int VEC>
__device__ __forceinline__ void ld_words(const uint8_t* p, uint32_t (&w)[VEC])
{
#ifdef IBLDPC_LD_NOALLOC
    // A/B experiment: message rows are read once per phase; keep them out of the L1 data array the look-up tables
    // share their banks with
    if constexpr (VEC == 4) {
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p));
    } else if constexpr (VEC == 2) {
        asm volatile("ld.global.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
    } else {
        asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(w[0]) : "l"(p));
    }
    return;
#endif
    if constexpr (VEC == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else if constexpr (VEC == 2) {
        const uint2 v = *reinterpret_cast<const uint2*>(p);
        w[0] = v.x; w[1] = v.y;
    } else {
        w[0] = *reinterpret_cast<const uint32_t*>(p);
    }
}
template <int VEC>
__device__ __forceinline__ void st_words(uint8_t* p, const uint32_t (&w)[VEC])
{
    if constexpr (VEC == 4) *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
    else *reinterpret_cast<uint32_t*>(p) = w[0];
}

// nibble f of w, multiplied by MUL (compile-time).  For a power-of-two MUL this is one shift and
// one mask (the mask fuses with the following OR/ADD into a single LOP3 / IADD3).
__host__ __device__ constexpr int n4_ilog2(uint32_t x) { return x <= 1 ? 0 : 1 + n4_ilog2(x >> 1); }

template <uint32_t MUL>
__device__ __forceinline__ uint32_t nib_times(uint32_t w, int f)
{
    if constexpr ((MUL & (MUL - 1)) == 0) {
        constexpr int sh = n4_ilog2(MUL);
        const int s = 4 * f - sh;
        const uint32_t x = s >= 0 ? (w >> s) : (w << (-s));
        return x & (15u << sh);
    } else {
        return ((w >> (4 * f)) & 15u) * MUL;
    }
}

// ------------------------------------------------------------------------------------------
// table staging: compact tables of this launch -> scratch (coalesced) -> lane-striped rows
// Host side: dynamic smem = n4_table_bytes(W) + stage_scratch_bytes(nst, T, dmax_match).
// Message alignment is folded into the last stage exactly as in stage_tables().
// ------------------------------------------------------------------------------------------
template <int W, int NT = kThreads>
__device__ __forceinline__ void stage_tables_n4(uint32_t* s_tab, const IbArgs& a, const uint8_t* lut)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int T = a.T, TT = T * T;
    constexpr int total = kTS * kTS * W;
    uint8_t* scratch = reinterpret_cast<uint8_t*>(s_tab) + (size_t)total * 128;
    const int n_lut = a.nst * TT;
    if (((n_lut | (int)(reinterpret_cast<uintptr_t>(lut) & 3)) & 3) == 0) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(lut);
        uint32_t* dst = reinterpret_cast<uint32_t*>(scratch);
        for (int i = threadIdx.x; i < n_lut / 4; i += NT) dst[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < n_lut; i += NT) scratch[i] = lut[i];
    }
    uint8_t* smatch = scratch + n_lut;
    if (a.match != nullptr)
        for (int i = threadIdx.x; i < a.dmax_match * T; i += NT) smatch[i] = a.match[i];
    __syncthreads();
    const bool fold = a.match != nullptr && a.nst >= 1;
    for (int rw = warp; rw < total; rw += NT / 32) {
        const int r = rw / W, w = rw - r * W;
#ifdef IBLDPC_DP4A
        const int t = r / kTS, m = r - t * kTS;   // row = t*16 + m: the message index is the minor one (stride 128 W bytes)
#else
        const int m = r / kTS, t = r - m * kTS;
#endif
        uint32_t v = 0;
        if (t < T) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = 4 * w + q;
                uint32_t e = 0;
                if (col < a.nst) {
                    if (m < T) {
                        e = scratch[col * TT + t * T + m];
                        if (fold && col == a.nst - 1) e = smatch[(a.dmax_match - 1) * T + e];
                        if (col == a.xp_col) e *= 4u;   // "x4" form: the value is the nibble shift into a tail-pair row
                    }
                } else if (col == a.nst && a.match != nullptr && !fold) {
                    if (m < a.dmax_match) e = smatch[m * T + t];
                }
                v |= e << (8 * q);
            }
        }
        s_tab[rw * 32 + lane] = v;
    }
}

#ifdef IBLDPC_DP4A
// ------------------------------------------------------------------------------------------
// dp4a address arithmetic (table rows t*16 + m: addr = t*TRS + m*RS + lane*4 + column offset, TRS = 16 RS).
// The eight nibbles of a message word are spread to two words of four BYTES once per word (even / odd frames); the
// address term of frame f is then ONE IDP.4A on the fma pipe -- byte (f >> 1) times the row stride plus the lane offset
// -- instead of a shift and a LOP3 on the alu pipe (the pipe ncu shows as the busiest: math-pipe throttle is the top
// stall of the check-node kernel).  Values and look-up order are unchanged: results stay bit-identical.
// ------------------------------------------------------------------------------------------
struct NibBytes { uint32_t e, o; };   // frames 0,2,4,6 and 1,3,5,7 of a word, one byte each
template <uint32_t W>
__device__ __forceinline__ NibBytes nib_bytes(uint32_t w)
{
    NibBytes b{w & 0x0f0f0f0fu, (w >> 4) & 0x0f0f0f0fu};
    if constexpr (W > 1) { b.e *= W; b.o *= W; }   // 15 W <= 255: no carry between the bytes
    return b;
}
// (byte of frame f) * MUL + add
template <uint32_t MUL>
__device__ __forceinline__ uint32_t byte_mad(const NibBytes& b, int f, uint32_t add)
{
    static_assert(MUL <= 255u, "dp4a multiplier is one byte");
    return __dp4a((f & 1) ? b.o : b.e, MUL << (8 * (f >> 1)), add);
}
// the (m_a << 4 | m_b) bytes of two message words: row index of the tail-pair table
__device__ __forceinline__ NibBytes pair_bytes(uint32_t wa, uint32_t wb)
{
    return NibBytes{((wa << 4) & 0xf0f0f0f0u) | (wb & 0x0f0f0f0fu), (wa & 0xf0f0f0f0u) | ((wb >> 4) & 0x0f0f0f0fu)};
}
#endif

// Output nibble of frame f selected from a composed 64-bit row: o collects frame f in nibble f.
// Default: push from the top -- o = (o >> 4) | (x << 28) is ONE funnel shift that also drops the bits of x above the
// nibble, and after the eight frames of a word (f = 0..7 in order) frame f sits in nibble f.  Two instructions per
// output and frame (SHF.R.U64 + SHF.R.W) instead of three (SHF, LOP3, IMAD/LEA); the kernels are issue-bound.
// Measured on B200: (3,6) check-node phase 0.423 -> 0.414 ms, DVB-S2 d_c = 7 0.584 -> 0.572 ms.  The translation units
// built with the shift / LOP3 address arithmetic (IBLDPC_NO_DP4A: the tail-pair variable-node kernels, alu-pipe bound)
// keep the mask-shift-add form -- there the extra SHF made the DVB-S2 variable-node phase slower (0.641 -> 0.667 ms).
// -DIBLDPC_NO_NIBPUSH: the mask-shift-add form everywhere (A/B builds).
__device__ __forceinline__ void nib_push(uint32_t& o, unsigned long long g, uint32_t e, int f)
{
#if defined(IBLDPC_NO_NIBPUSH) || !defined(IBLDPC_DP4A)
    o += ((uint32_t)(g >> e) & 15u) << (4 * f);
#else
    (void)f;
    o = __funnelshift_r(o, (uint32_t)(g >> e), 4);
#endif
}

// ------------------------------------------------------------------------------------------
// check node, 8 frames (one 32-bit word per message)
// ------------------------------------------------------------------------------------------
// WT / CB: words per table row and first stage column of this degree class inside a table image shared by several
// classes (fused per-phase kernels, ib_phase_n4.cuh); WT = 0 selects the class's own layout (column 0, n4_*_words).
template <int D, bool MATCH, int WT = 0, int CB = 0>
__device__ __forceinline__ void cn_word_n4(const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab, uint32_t lane4)
{
    constexpr uint32_t W = WT ? WT : n4_cn_words(D, MATCH), RS = 128u * W, TRS = RS * kTS;
#ifdef IBLDPC_DP4A
    const uint32_t match_off = (uint32_t)(D - 1) * RS + IB_SO(CB + D - 2) + lane4;
    NibBytes b[D], r0 = nib_bytes<1>(w[0]), r1 = nib_bytes<1>(w[1]);
#pragma unroll
    for (int k = 0; k < D; ++k) b[k] = nib_bytes<W>(w[k]);
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        uint32_t ms[D];
#pragma unroll
        for (int k = 1; k < D; ++k) ms[k] = byte_mad<128u>(b[k], f, lane4);
        uint32_t P[D > 1 ? D : 2];
        P[1] = byte_mad<1u>(r0, f, 0u);
#pragma unroll
        for (int j = 1; j <= D - 2; ++j) P[j + 1] = lut_ld(tab, P[j] * TRS + ms[j] + IB_SO(CB + j - 1));
#pragma unroll
        for (int wo = 0; wo < D; ++wo) {
            uint32_t t = (wo == 0) ? byte_mad<1u>(r1, f, 0u) : P[wo];
#pragma unroll
            for (int k = (wo == 0 ? 2 : wo + 1); k < D; ++k) t = lut_ld(tab, t * TRS + ms[k] + IB_SO(CB + k - 2));
            if (MATCH) t = lut_ld(tab, t * TRS + match_off);
            o[wo] += t << (4 * f);
        }
    }
#else
    const uint32_t match_off = (uint32_t)(D - 1) * TRS + IB_SO(CB + D - 2) + lane4;
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        uint32_t ms[D];
#pragma unroll
        for (int k = 0; k < D; ++k) ms[k] = nib_times<TRS>(w[k], f) | lane4;
        uint32_t P[D > 1 ? D : 2];
        P[1] = (w[0] >> (4 * f)) & 15u;
#pragma unroll
        for (int j = 1; j <= D - 2; ++j) P[j + 1] = lut_ld(tab, P[j] * RS + ms[j] + IB_SO(CB + j - 1));
#pragma unroll
        for (int wo = 0; wo < D; ++wo) {
            uint32_t t = (wo == 0) ? ((w[1] >> (4 * f)) & 15u) : P[wo];
#pragma unroll
            for (int k = (wo == 0 ? 2 : wo + 1); k < D; ++k) t = lut_ld(tab, t * RS + ms[k] + IB_SO(CB + k - 2));
            if (MATCH) t = lut_ld(tab, t * RS + match_off);
            o[wo] += t << (4 * f);
        }
    }
#endif
}

// Tail-pair variant (D >= 4), see cn_word_pair in ib_kernels.cuh: all outputs w <= D-3 end with
//   out_w = S_{D-3}( S_{D-4}(x_w, m_{D-2}), m_{D-1} ) = G(m_{D-2}, m_{D-1})[x_w],
// G composed on the host (ibldpc_set_luts), 16 nibbles = one 64-bit row per (m_{D-2}, m_{D-1}),
// fetched with ONE conflict-free LDS.64 per frame (kPairSlots lane slots).  The stage feeding the
// row (column D-5) is stored as 4*x, so selecting nibble x_w is a 64-bit shift and a mask.
// Shared-memory wavefronts per check and frame: D=6 18 -> 12, D=7 25 -> 16, D=8 33 -> 21.
constexpr uint32_t kPairBytes = kTS * kTS * 8 * kPairSlots;   // 32 KB, placed in front of the stage tables

template <int D, int WT = 0, int CB = 0>
__device__ __forceinline__ void cn_word_n4_pair(const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab,
                                                const uint8_t* ptab, uint32_t lane4, uint32_t slot8)
{
    static_assert(D >= 4, "tail-pair variant needs at least two look-up stages");
    constexpr uint32_t W = WT ? WT : n4_cn_words(D, false), RS = 128u * W, TRS = RS * kTS;
    [[maybe_unused]] constexpr uint32_t PS = 8u * kPairSlots, TPS = PS * kTS;
#ifdef IBLDPC_DP4A
    static_assert(PS <= 255u, "pair row stride must fit the dp4a multiplier");
    NibBytes b[D], r0 = nib_bytes<1>(w[0]), r1 = nib_bytes<1>(w[1]);
#pragma unroll
    for (int k = 1; k < D; ++k) b[k] = nib_bytes<W>(w[k]);
    const NibBytes pb = pair_bytes(w[D - 2], w[D - 1]);
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        uint32_t ms[D];
#pragma unroll
        for (int k = 1; k < D; ++k) ms[k] = byte_mad<128u>(b[k], f, lane4);
        const uint2 g2 = *reinterpret_cast<const uint2*>(ptab + byte_mad<PS>(pb, f, slot8));
        const unsigned long long g = ((unsigned long long)g2.y << 32) | g2.x;
        // prefix chain; P[D-3] comes out of column D-5, i.e. in x4 form (D >= 5)
        uint32_t P[D];
        P[1] = byte_mad<1u>(r0, f, 0u);
#pragma unroll
        for (int j = 1; j <= D - 3; ++j) {
            const bool in_x4 = (D >= 5) && (j == D - 3);
            P[j + 1] = lut_ld(tab, P[j] * (in_x4 ? TRS / 4u : TRS) + ms[j] + IB_SO(CB + j - 1));
        }
        // the two outputs that skip one of the tail messages
        o[D - 1] += lut_ld(tab, P[D - 2] * TRS + ms[D - 2] + IB_SO(CB + D - 3)) << (4 * f);
        o[D - 2] += lut_ld(tab, P[D - 2] * TRS + ms[D - 1] + IB_SO(CB + D - 3)) << (4 * f);
#pragma unroll
        for (int wo = 0; wo <= D - 3; ++wo) {
            uint32_t e;   // 4 * x_w
            if (D >= 5 && wo == D - 3) {
                e = P[D - 3];
            } else if (D == 4) {
                e = byte_mad<4u>(wo == 0 ? r1 : r0, f, 0u);
            } else {
                uint32_t t = (wo == 0) ? byte_mad<1u>(r1, f, 0u) : P[wo];
#pragma unroll
                for (int k = (wo == 0 ? 2 : wo + 1); k <= D - 3; ++k) t = lut_ld(tab, t * TRS + ms[k] + IB_SO(CB + k - 2));
                e = t;   // the last look-up read column D-5
            }
            nib_push(o[wo], g, e, f);
        }
    }
#else
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        uint32_t ms[D];
#pragma unroll
        for (int k = 1; k < D; ++k) ms[k] = nib_times<TRS>(w[k], f) | lane4;
        const uint2 g2 = *reinterpret_cast<const uint2*>(ptab + (nib_times<TPS>(w[D - 2], f) | nib_times<PS>(w[D - 1], f) | slot8));
        const unsigned long long g = ((unsigned long long)g2.y << 32) | g2.x;
        // prefix chain; P[D-3] comes out of column D-5, i.e. in x4 form (D >= 5)
        uint32_t P[D];
        P[1] = (w[0] >> (4 * f)) & 15u;
#pragma unroll
        for (int j = 1; j <= D - 3; ++j) {
            const bool in_x4 = (D >= 5) && (j == D - 3);
            P[j + 1] = lut_ld(tab, P[j] * (in_x4 ? RS / 4u : RS) + ms[j] + IB_SO(CB + j - 1));
        }
        // the two outputs that skip one of the tail messages
        o[D - 1] += lut_ld(tab, P[D - 2] * RS + ms[D - 2] + IB_SO(CB + D - 3)) << (4 * f);
        o[D - 2] += lut_ld(tab, P[D - 2] * RS + ms[D - 1] + IB_SO(CB + D - 3)) << (4 * f);
#pragma unroll
        for (int wo = 0; wo <= D - 3; ++wo) {
            uint32_t e;   // 4 * x_w
            if (D >= 5 && wo == D - 3) {
                e = P[D - 3];
            } else if (D == 4) {
                e = nib_times<4u>(w[wo == 0 ? 1 : 0], f);
            } else {
                uint32_t t = (wo == 0) ? ((w[1] >> (4 * f)) & 15u) : P[wo];
#pragma unroll
                for (int k = (wo == 0 ? 2 : wo + 1); k <= D - 3; ++k) t = lut_ld(tab, t * RS + ms[k] + IB_SO(CB + k - 2));
                e = t;   // the last look-up read column D-5
            }
            nib_push(o[wo], g, e, f);
        }
    }
#endif
}

// FS: per-frame syndrome flags (per-frame early termination, ib_perframe.cu): 0 none, 1 OR-ed into a shared-memory
// accumulator at the 32-bit shared address fsyn_s, 2 OR-ed into global memory at fsyn.  Both are issued UNCONDITIONALLY: a
// branch around them in this fully unrolled body makes ptxas keep the shared-memory base of the tables in a register and
// add it to every look-up address (one extra IMAD.IADD per LDS: +23 % instructions, measured in round 2).
// degree-6 check nodes through the three-input table of the first two stages (ib_triple_n4.cuh; TRI = 1 below)
__device__ __forceinline__ void cn6_word_n4_triple(const uint32_t (&w)[6], uint32_t (&o)[6], const uint8_t* tab, const uint8_t* ptab,
                                                   const uint8_t* ttab, uint32_t lane4, uint32_t slot8);
template <int D>   // degrees 7 and 8 (local stage columns: stage column c + 2 is column c of the staged table word)
__device__ __forceinline__ void cn_word_n4_triple(const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab, const uint8_t* ptab,
                                                  const uint8_t* ttab, uint32_t lane4, uint32_t slot8);

template <int D, bool MATCH, bool EARLY, int VEC, bool PAIR, int WT = 0, int CB = 0, int FS = 0, int TRI = 0>
__device__ __forceinline__ uint32_t cn_node_n4(const IbArgs& a, const uint8_t* tab, const uint8_t* ptab, int s, uint32_t col,
                                               uint32_t lane4, int valid_frames, uint32_t* fsyn = nullptr,
                                               const uint32_t* frz = nullptr, uint32_t fsyn_s = 0xffffffffu)
{
    uint32_t m[D][VEC];
    if (a.iter0) {
#pragma unroll
        for (int k = 0; k < D; ++k) ld_words<VEC>(a.ch + (uint64_t)(uint32_t)a.vidx[s + k] * a.pitch + col, m[k]);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) ld_words<VEC>(a.msg + (uint64_t)(uint32_t)(s + k) * a.pitch + col, m[k]);
    }
    uint32_t syn = 0;
    uint32_t r[D][VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        uint32_t w[D], o[D];
#pragma unroll
        for (int k = 0; k < D; ++k) w[k] = m[k][j];
        if (EARLY && !a.iter0) {
            // calc_syndrome (kernels_template_irreg.cl:304-325) on the VN->CN messages just read:
            // parity of (msg < T/2) over the D inputs, one bit per frame nibble
            uint32_t par = 0;
            if (a.tshift >= 0) {   // T power of two: (m < T/2) == !bit(log2(T)-1)
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) x ^= w[k];
                par = ((x >> a.tshift) & 0x11111111u) ^ ((D & 1) ? 0x11111111u : 0u);
            } else {
#pragma unroll
                for (int f = 0; f < 8; ++f) {
                    uint32_t p1 = 0;
#pragma unroll
                    for (int k = 0; k < D; ++k) p1 ^= (((w[k] >> (4 * f)) & 15u) < (uint32_t)(a.T / 2)) ? 1u : 0u;
                    par |= p1 << (4 * f);
                }
            }
            const int nv = valid_frames - 8 * j;   // ignore padding frames
            const uint32_t vmask = nv >= 8 ? 0xffffffffu : nv <= 0 ? 0u : ((1u << (4 * nv)) - 1u);
            syn |= par & vmask;
            // per-frame early termination: bit 4f of fsyn[word] = frame f of that word failed a check in this pass
            if constexpr (FS == 1) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(fsyn_s + 4u * j), "r"(par & vmask));
            if constexpr (FS == 2) asm volatile("red.global.or.b32 [%0], %1;" ::"l"(fsyn + j), "r"(par & vmask));
        }
        if constexpr (TRI != 0 && D == 6) cn6_word_n4_triple(w, o, tab, ptab, ptab + kPairBytes, lane4, (lane4 & (4u * (kPairSlots - 1))) * 2u);
        else if constexpr (TRI != 0) cn_word_n4_triple<D>(w, o, tab, ptab, ptab + kPairBytes, lane4, (lane4 & (4u * (kPairSlots - 1))) * 2u);
        else if constexpr (PAIR) cn_word_n4_pair<D, WT, CB>(w, o, tab, ptab, lane4, (lane4 & (4u * (kPairSlots - 1))) * 2u);
        else cn_word_n4<D, MATCH, WT, CB>(w, o, tab, lane4);
        // per-frame early termination: frames that have converged keep the messages they converged with (nibble mask
        // frz[j] = frames of word j that still iterate); they are decided later from exactly this state (ib_perframe.cu)
        if (frz != nullptr && !a.iter0) {
            const uint32_t am = frz[j];
#pragma unroll
            for (int k = 0; k < D; ++k) o[k] = (o[k] & am) | (w[k] & ~am);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) r[k][j] = o[k];
    }
#pragma unroll
    for (int k = 0; k < D; ++k) st_words<VEC>(a.msg + (uint64_t)(uint32_t)(s + k) * a.pitch + col, r[k]);
    return syn;
}

__host__ __device__ constexpr int cn_n4_min_blocks(int D, int VEC, bool PAIR)
{
    return PAIR ? (D <= 6 ? 3 : 2) : VEC == 2 ? (D <= 6 ? 5 : (D <= 8 ? 3 : 2)) : (D <= 6 ? 4 : (D <= 8 ? 3 : 2));
}

// Node loop of one check-node launch (or of one check-node phase of the cooperative kernel): CTA -> (tile group
// blockIdx.y, node subset), no divisions, 32-bit offsets; the next node's slot index is fetched while this node is
// being computed.  Returns the OR of the syndrome bits this thread saw.
template <int D, bool MATCH, bool EARLY, int VEC, bool PAIR, int NT, int TRI = 0>
__device__ __forceinline__ uint32_t cn_loop_n4(const IbArgs& a, const uint8_t* tab, const uint8_t* ptab,
                                               const int* __restrict__ nodes, int n_nodes)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane4 = lane * 4;
    const int tile = (blockIdx.y << a.tpc_log2) + (warp & ((1 << a.tpc_log2) - 1));
    const int nps = (NT / 32) >> a.tpc_log2;
    const int stride = gridDim.x * nps;
    const uint32_t col = ((uint32_t)tile * 32u + lane) * (4u * VEC);
    uint32_t syn = 0;
    if (tile < a.tiles && col < a.pitch) {
        const int valid = a.B - 2 * (int)col;
        int i = blockIdx.x * nps + (warp >> a.tpc_log2);
        int s = i < n_nodes ? a.sc[nodes[i]] : 0;
        while (i < n_nodes) {
            const int i2 = i + stride;
            const int s2 = i2 < n_nodes ? a.sc[nodes[i2]] : 0;
            syn |= cn_node_n4<D, MATCH, EARLY, VEC, PAIR, 0, 0, 0, TRI>(a, tab, ptab, s, col, lane4, valid);
            i = i2;
            s = s2;
        }
    }
    return syn;
}

// send + checknode_update_iter0 (a.iter0) or checknode_update + calc_syndrome, packed nibbles.
// PAIR: shared memory = [tail-pair rows (kPairBytes)][stage tables][staging scratch].
// NT = threads per CTA: the tail-pair kernels of degree <= 8 run 512 threads (2 CTAs/SM at 64 registers share
// one 64-96 KB table set per 16 warps: 32 warps/SM instead of 24 resp. 16 with 256-thread CTAs).
template <int D, bool MATCH, bool EARLY, int VEC, bool PAIR, int NT = kThreads>
__global__ void __launch_bounds__(NT, NT == 1024 ? 1 : NT == 512 ? 2 : cn_n4_min_blocks(D, VEC, PAIR))
ib_cn_n4_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_all[];
    if (EARLY && a.it >= 1 && a.flags[a.it - 1] == 0) return;   // batch already converged
    uint32_t* s_tab = s_all + (PAIR ? kPairBytes / 4 : 0);
    if (PAIR) {
        // expand the composed rows (global: row a*T+b, 8 bytes) to kPairSlots lane slots per row a*16+b
        const uint2* src = reinterpret_cast<const uint2*>(a.pair);
        uint2* dst = reinterpret_cast<uint2*>(s_all);
        for (int i = threadIdx.x; i < kTS * kTS * kPairSlots; i += NT) {
            const int r = i / kPairSlots, ra = r / kTS, rb = r - ra * kTS;
            dst[i] = (ra < a.T && rb < a.T) ? src[ra * a.T + rb] : make_uint2(0u, 0u);
        }
    }
    stage_tables_n4<n4_cn_words(D, MATCH), NT>(s_tab, a, a.lut);
    __syncthreads();
    const uint8_t* tab = reinterpret_cast<const uint8_t*>(s_tab);
    const uint8_t* ptab = reinterpret_cast<const uint8_t*>(s_all);
    const uint32_t syn = cn_loop_n4<D, MATCH, EARLY, VEC, PAIR, NT>(a, tab, ptab, nodes, n_nodes);
    if (EARLY && !a.iter0) {
        const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
        if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[a.it], 1);
    }
}

// ------------------------------------------------------------------------------------------
// variable node: channel value + D inbox messages, 8 frames
// ------------------------------------------------------------------------------------------
template <int D, bool DECIDE, int WT = 0, int CB = 0>
__device__ __forceinline__ void vn_word_n4(uint32_t chw, const uint32_t (&w)[D], uint32_t (&o)[D], uint32_t& dec_lo,
                                           uint32_t& dec_hi, const uint8_t* tab, uint32_t lane4)
{
    constexpr uint32_t W = WT ? WT : n4_vn_words(D, DECIDE), RS = 128u * W, TRS = RS * kTS;
#ifdef IBLDPC_DP4A
    constexpr uint32_t TSTR = TRS;   // stride of the running value t (rows t*16 + m)
    NibBytes b[D + 1];
    const NibBytes rc = nib_bytes<1>(chw);
#pragma unroll
    for (int k = 1; k <= D; ++k) b[k] = nib_bytes<W>(w[k - 1]);
#else
    constexpr uint32_t TSTR = RS;
#endif
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
    dec_lo = dec_hi = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        uint32_t ms[D + 1];   // ms[k] for y_k, k = 1..D
        uint32_t P[D + 2];
#ifdef IBLDPC_DP4A
#pragma unroll
        for (int k = 1; k <= D; ++k) ms[k] = byte_mad<128u>(b[k], f, lane4);
        P[1] = byte_mad<1u>(rc, f, 0u);
#else
#pragma unroll
        for (int k = 1; k <= D; ++k) ms[k] = nib_times<TRS>(w[k - 1], f) | lane4;
        P[1] = (chw >> (4 * f)) & 15u;
#endif
#pragma unroll
        for (int j = 1; j <= D - 1; ++j) P[j + 1] = lut_ld(tab, P[j] * TSTR + ms[j] + IB_SO(CB + j - 1));
        if (DECIDE) {
            const uint32_t t = lut_ld(tab, P[D] * TSTR + ms[D] + IB_SO(CB + D - 1));
            if (f < 4) dec_lo = put_byte(dec_lo, t, f & 3);
            else dec_hi = put_byte(dec_hi, t, f & 3);
        } else {
#pragma unroll
            for (int wo = 1; wo <= D; ++wo) {
                uint32_t t = P[wo];
#pragma unroll
                for (int k = wo + 1; k <= D; ++k) t = lut_ld(tab, t * TSTR + ms[k] + IB_SO(CB + k - 2));
                o[wo - 1] += t << (4 * f);
            }
        }
    }
}

// Tail-pair variant of the variable-node update (D >= 3): all outputs w <= D-2 end with
//   out_w = S_{D-2}( S_{D-3}(x_w, y_{D-1}), y_D ) = G(y_{D-1}, y_D)[x_w]
// (G composed on the host with the matching row of degree D folded in).  One LDS.64 per frame
// replaces 2(D-2) look-ups; column D-4, which produces every x_w, is stored as 4*x.
// Shared-memory wavefronts per variable node and frame: D=8 35 -> 25, D=11 65 -> 49.
template <int D, int WT = 0, int CB = 0>
__device__ __forceinline__ void vn_word_n4_pair(uint32_t chw, const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab,
                                                const uint8_t* ptab, uint32_t lane4, uint32_t slot8)
{
    static_assert(D >= 3, "tail-pair variant needs two update stages");
    constexpr uint32_t W = WT ? WT : n4_vn_words(D, false), RS = 128u * W, TRS = RS * kTS;
    [[maybe_unused]] constexpr uint32_t PS = 8u * kPairSlots, TPS = PS * kTS;
#ifdef IBLDPC_DP4A
    constexpr uint32_t TSTR = TRS;
    NibBytes b[D + 1];
    const NibBytes rc = nib_bytes<1>(chw);
#pragma unroll
    for (int k = 1; k <= D; ++k) b[k] = nib_bytes<W>(w[k - 1]);
    const NibBytes pb = pair_bytes(w[D - 2], w[D - 1]);
#else
    constexpr uint32_t TSTR = RS;
#endif
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        uint32_t ms[D + 1];   // ms[k] for y_k, k = 1..D
        uint32_t P[D + 1];
#ifdef IBLDPC_DP4A
#pragma unroll
        for (int k = 1; k <= D; ++k) ms[k] = byte_mad<128u>(b[k], f, lane4);
        const uint2 g2 = *reinterpret_cast<const uint2*>(ptab + byte_mad<PS>(pb, f, slot8));
        P[1] = byte_mad<1u>(rc, f, 0u);
#else
#pragma unroll
        for (int k = 1; k <= D; ++k) ms[k] = nib_times<TRS>(w[k - 1], f) | lane4;
        const uint2 g2 = *reinterpret_cast<const uint2*>(ptab + (nib_times<TPS>(w[D - 2], f) | nib_times<PS>(w[D - 1], f) | slot8));
        P[1] = (chw >> (4 * f)) & 15u;
#endif
        const unsigned long long g = ((unsigned long long)g2.y << 32) | g2.x;
        // prefix chain P[1..D-1]; P[D-2] comes out of column D-4, i.e. as 4*x (D >= 4)
#pragma unroll
        for (int j = 1; j <= D - 2; ++j) {
            const bool in_x4 = (D >= 4) && (j == D - 2);
            P[j + 1] = lut_ld(tab, P[j] * (in_x4 ? TSTR / 4u : TSTR) + ms[j] + IB_SO(CB + j - 1));
        }
        // the two outputs that skip one of the tail messages
        o[D - 2] += lut_ld(tab, P[D - 1] * TSTR + ms[D] + IB_SO(CB + D - 2)) << (4 * f);       // w = D-1
        o[D - 1] += lut_ld(tab, P[D - 1] * TSTR + ms[D - 1] + IB_SO(CB + D - 2)) << (4 * f);   // w = D
#pragma unroll
        for (int wo = 1; wo <= D - 2; ++wo) {
            uint32_t e;   // 4 * x_w
            if (D >= 4 && wo == D - 2) {
                e = P[D - 2];
            } else if (D == 3) {
#ifdef IBLDPC_DP4A
                e = byte_mad<4u>(rc, f, 0u);
#else
                e = nib_times<4u>(chw, f);
#endif
            } else {
                uint32_t t = P[wo];
#pragma unroll
                for (int k = wo + 1; k <= D - 2; ++k) t = lut_ld(tab, t * TSTR + ms[k] + IB_SO(CB + k - 2));
                e = t;   // the last look-up read column D-4
            }
            nib_push(o[wo - 1], g, e, f);
        }
    }
}

template <int D, int VEC> struct VnIn4 { uint32_t c[VEC]; uint32_t m[D][VEC]; };

template <int D, int VEC>
__device__ __forceinline__ void vn_load_msgs_n4(const IbArgs& a, const VnIdx<D>& x, uint32_t col, VnIn4<D, VEC>& in)
{
    if (x.ok) {
        ld_words<VEC>(a.ch + (uint64_t)(uint32_t)x.v * a.pitch + col, in.c);
#pragma unroll
        for (int k = 0; k < D; ++k) ld_words<VEC>(a.msg + (uint64_t)(uint32_t)x.rows[k] * a.pitch + col, in.m[k]);
    }
}

template <int D, bool DECIDE, int VEC, bool PAIR = false>
__device__ __forceinline__ void vn_compute_store_n4(const IbArgs& a, const uint8_t* tab, const uint8_t* ptab, const VnIdx<D>& x,
                                                    const VnIn4<D, VEC>& in, uint32_t col, uint32_t lane4)
{
    if (!DECIDE && D == 1) {   // degree-1 variable node forwards the raw channel value (:132-136)
        st_words<VEC>(a.msg + (uint64_t)(uint32_t)x.rows[0] * a.pitch + col, in.c);
        return;
    }
    uint32_t r[D][VEC];
    uint32_t dec[2 * VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        uint32_t w[D], o[D];
#pragma unroll
        for (int k = 0; k < D; ++k) w[k] = in.m[k][j];
        if constexpr (PAIR) vn_word_n4_pair<D>(in.c[j], w, o, tab, ptab, lane4, (lane4 & (4u * (kPairSlots - 1))) * 2u);
        else vn_word_n4<D, DECIDE>(in.c[j], w, o, dec[2 * j], dec[2 * j + 1], tab, lane4);
        if (!DECIDE) {
#pragma unroll
            for (int k = 0; k < D; ++k) r[k][j] = o[k];
        }
    }
    if (DECIDE) {
        // decided cluster indices leave as uint8 (one byte per frame): 8*VEC bytes per lane
        const uint32_t ocol = 2u * col;
        uint8_t* dst = a.out + (uint64_t)(uint32_t)x.v * a.out_pitch + ocol;
#pragma unroll
        for (int q = 0; q < VEC / 2; ++q)
            if (ocol + 16u * q < a.out_pitch)
                *reinterpret_cast<uint4*>(dst + 16 * q) = make_uint4(dec[4 * q], dec[4 * q + 1], dec[4 * q + 2], dec[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) st_words<VEC>(a.msg + (uint64_t)(uint32_t)x.rows[k] * a.pitch + col, r[k]);
    }
}

// Software pipeline over the node list of one degree class (row indices fetched two nodes ahead,
// see vn_loop in ib_kernels.cuh).
template <int D, bool DECIDE, int VEC, bool PAIR = false, int NT = kThreads>
__device__ __forceinline__ void vn_loop_n4(const IbArgs& a, const uint8_t* tab, const uint8_t* ptab, const int* __restrict__ nodes,
                                           int n_nodes)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane4 = lane * 4;
    const int tile = (blockIdx.y << a.tpc_log2) + (warp & ((1 << a.tpc_log2) - 1));
    const int nps = (NT / 32) >> a.tpc_log2;
    const int stride = gridDim.x * nps;
    const uint32_t col = ((uint32_t)tile * 32u + lane) * (4u * VEC);
    if (tile >= a.tiles || col >= a.pitch) return;
    int i = blockIdx.x * nps + (warp >> a.tpc_log2);
    VnIdx<D> cur, nxt, nn;
    vn_load_idx<D>(a, nodes, n_nodes, i, cur);
    vn_load_idx<D>(a, nodes, n_nodes, i + stride, nxt);
    int inn = i + 2 * stride;
    while (cur.ok) {
        VnIn4<D, VEC> buf;
        vn_load_msgs_n4<D, VEC>(a, cur, col, buf);
        vn_load_idx<D>(a, nodes, n_nodes, inn, nn);
        inn += stride;
        vn_compute_store_n4<D, DECIDE, VEC, PAIR>(a, tab, ptab, cur, buf, col, lane4);
        cur = nxt;
        nxt = nn;
    }
}

// ------------------------------------------------------------------------------------------
// Small batches (B <= 256: the reference's DVB-S2 / WLAN drivers decode msg_at_time = 2 frames per call,
// Irregular_LDPC_Decoding/DVB-S2/BER_simulation_OpenCL.py:71): with one warp per (node, tile) all but one or two lanes
// of a warp have no frames and a warp walks through its nodes one after the other.  Here a LANE is a (node, word) pair:
// wpn = ceil(B / 8) words per node, 32 / wpn nodes per warp step.  The table copies are per lane, so the look-up
// functions (cn_word_n4 / cn_word_n4_pair / vn_word_n4) are used unchanged: same chains, bit-identical results.
// ------------------------------------------------------------------------------------------
constexpr int kLaneModeMaxFrames = 256;   // up to 32 words per node: never fewer active lanes than the (node, tile) mapping

template <int D, bool MATCH, bool EARLY, bool PAIR, int NT, int WT = 0, int CB = 0>
__device__ __forceinline__ uint32_t cn_lanes_n4(const IbArgs& a, const uint8_t* tab, const uint8_t* ptab,
                                                const int* __restrict__ nodes, int n_nodes)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lane4 = lane * 4;
    const int wpn = (a.B + 7) >> 3, per_warp = 32 / wpn;
    const int j = lane / wpn, wd = lane - j * wpn;
    const long long gw = (long long)(blockIdx.x + blockIdx.y * gridDim.x) * (NT / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * gridDim.y * (NT / 32);
    uint32_t syn = 0;
    for (long long base = gw * per_warp; base < n_nodes; base += nw * per_warp) {
        const int i = (int)base + j;
        if (j >= per_warp || i >= n_nodes) continue;
        const int s = a.sc[nodes[i]];
        uint32_t w[D], o[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const uint8_t* row = a.iter0 ? a.ch + (uint64_t)(uint32_t)a.vidx[s + k] * a.pitch : a.msg + (uint64_t)(uint32_t)(s + k) * a.pitch;
            w[k] = *reinterpret_cast<const uint32_t*>(row + 4 * wd);
        }
        if (EARLY && !a.iter0) {   // calc_syndrome on the VN->CN messages just read, see cn_node_n4
            uint32_t par = 0;
            if (a.tshift >= 0) {
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) x ^= w[k];
                par = ((x >> a.tshift) & 0x11111111u) ^ ((D & 1) ? 0x11111111u : 0u);
            } else {
#pragma unroll
                for (int f = 0; f < 8; ++f) {
                    uint32_t p1 = 0;
#pragma unroll
                    for (int k = 0; k < D; ++k) p1 ^= (((w[k] >> (4 * f)) & 15u) < (uint32_t)(a.T / 2)) ? 1u : 0u;
                    par |= p1 << (4 * f);
                }
            }
            const int nv = a.B - 8 * wd;
            syn |= par & (nv >= 8 ? 0xffffffffu : ((1u << (4 * nv)) - 1u));
        }
        if constexpr (PAIR) cn_word_n4_pair<D, WT, CB>(w, o, tab, ptab, lane4, (lane4 & (4u * (kPairSlots - 1))) * 2u);
        else cn_word_n4<D, MATCH, WT, CB>(w, o, tab, lane4);
#pragma unroll
        for (int k = 0; k < D; ++k) *reinterpret_cast<uint32_t*>(a.msg + (uint64_t)(uint32_t)(s + k) * a.pitch + 4 * wd) = o[k];
    }
    return syn;
}

template <int D, bool DECIDE, int NT, int WT = 0, int CB = 0, bool PAIR = false>
__device__ __forceinline__ void vn_lanes_n4(const IbArgs& a, const uint8_t* tab, const int* __restrict__ nodes, int n_nodes,
                                            const uint8_t* ptab = nullptr)
{
    const int lane = threadIdx.x & 31;
    const uint32_t lane4 = lane * 4;
    const int wpn = (a.B + 7) >> 3, per_warp = 32 / wpn;
    const int j = lane / wpn, wd = lane - j * wpn;
    const long long gw = (long long)(blockIdx.x + blockIdx.y * gridDim.x) * (NT / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * gridDim.y * (NT / 32);
    for (long long base = gw * per_warp; base < n_nodes; base += nw * per_warp) {
        const int i = (int)base + j;
        if (j >= per_warp || i >= n_nodes) continue;
        const int v = nodes[i];
        const int s = a.sv[v];
        const uint32_t chw = *reinterpret_cast<const uint32_t*>(a.ch + (uint64_t)(uint32_t)v * a.pitch + 4 * wd);
        int rows[D];
        uint32_t w[D], o[D], dlo = 0, dhi = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
            rows[k] = a.tv[s + k];
            w[k] = *reinterpret_cast<const uint32_t*>(a.msg + (uint64_t)(uint32_t)rows[k] * a.pitch + 4 * wd);
        }
        if (!DECIDE && D == 1) {   // degree-1 variable node forwards the raw channel value (:132-136)
            *reinterpret_cast<uint32_t*>(a.msg + (uint64_t)(uint32_t)rows[0] * a.pitch + 4 * wd) = chw;
            continue;
        }
        if constexpr (PAIR && !DECIDE) vn_word_n4_pair<D, WT, CB>(chw, w, o, tab, ptab, lane4, (lane4 & (4u * (kPairSlots - 1))) * 2u);
        else vn_word_n4<D, DECIDE, WT, CB>(chw, w, o, dlo, dhi, tab, lane4);
        if (DECIDE) {
            if (8u * (uint32_t)wd < a.out_pitch)
                *reinterpret_cast<uint2*>(a.out + (uint64_t)(uint32_t)v * a.out_pitch + 8 * wd) = make_uint2(dlo, dhi);
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) *reinterpret_cast<uint32_t*>(a.msg + (uint64_t)(uint32_t)rows[k] * a.pitch + 4 * wd) = o[k];
        }
    }
}

__host__ __device__ constexpr int vn_n4_min_blocks(int D, int VEC)
{
    return VEC == 2 ? (D <= 4 ? 5 : (D <= 8 ? 3 : 2)) : (D <= 4 ? 4 : (D <= 8 ? 3 : 2));
}

// varnode_update (kernels_template_irreg.cl:103-179), packed nibbles
template <int D, int VEC, int NT = kThreads>
__global__ void __launch_bounds__(NT, NT == 1024 ? 1 : NT == 512 ? 2 : vn_n4_min_blocks(D, VEC))
ib_vn_n4_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;
    if (D > 1) {
        stage_tables_n4<n4_vn_words(D, false), NT>(s_tab, a, a.lut);
        __syncthreads();
    }
    vn_loop_n4<D, false, VEC, false, NT>(a, reinterpret_cast<const uint8_t*>(s_tab), nullptr, nodes, n_nodes);
}

// varnode_update through the tail-pair rows (vn_word_n4_pair): shared memory =
// [tail-pair rows (kPairBytes)][stage tables][staging scratch]; NT threads per CTA, 2 words per lane.
__host__ __device__ constexpr int vn_n4_pair_min_blocks(int D, int NT)
{
    return NT >= 640 ? 1 : NT == 512 ? (D <= 9 ? 2 : 1) : (D <= 5 ? 3 : (D <= 9 ? 2 : 1));
}
template <int D, int NT>
__global__ void __launch_bounds__(NT, vn_n4_pair_min_blocks(D, NT))
ib_vn_n4_pair_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_all[];
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;
    uint32_t* s_tab = s_all + kPairBytes / 4;
    const uint2* src = reinterpret_cast<const uint2*>(a.pair);
    uint2* dst = reinterpret_cast<uint2*>(s_all);
    for (int i = threadIdx.x; i < kTS * kTS * kPairSlots; i += NT) {
        const int r = i / kPairSlots, ra = r / kTS, rb = r - ra * kTS;
        dst[i] = (ra < a.T && rb < a.T) ? src[ra * a.T + rb] : make_uint2(0u, 0u);
    }
    stage_tables_n4<n4_vn_words(D, false), NT>(s_tab, a, a.lut);
    __syncthreads();
    vn_loop_n4<D, false, 2, true, NT>(a, reinterpret_cast<const uint8_t*>(s_tab), reinterpret_cast<const uint8_t*>(s_all), nodes,
                                      n_nodes);
}

// calc_varnode_output (kernels_template_irreg.cl:249-302) with the VN table of iteration i_num-1;
// reads packed nibbles, writes uint8 cluster indices
template <int D, int VEC>
__global__ void __launch_bounds__(kThreads) ib_out_n4_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    __shared__ int s_passes;
    if (threadIdx.x == 0) {
        s_passes = executed_passes(a);
        if (blockIdx.x == 0 && blockIdx.y == 0) *a.inum = s_passes + 1;
    }
    __syncthreads();
    stage_tables_n4<n4_vn_words(D, true)>(s_tab, a, a.lut + (long long)s_passes * a.vn_it_stride);
    __syncthreads();
    vn_loop_n4<D, true, VEC>(a, reinterpret_cast<const uint8_t*>(s_tab), nullptr, nodes, n_nodes);
}

// ------------------------------------------------------------------------------------------
// uint8 (rows, src_pitch) -> packed nibbles (rows, pitch4); frames >= B become 0.
// One thread per 32-bit destination word (8 frames).  Values >= T (not a cluster index of this decoder) are
// clamped to T-1 so that no look-up leaves its table, and reported through *bad (checked by the host at its next
// synchronisation point: ibldpc_last_i_num / the host-buffer entry points).
// ------------------------------------------------------------------------------------------
template <bool ALIGNED>
__global__ void pack_n4_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int rows, long long B,
                               long long src_pitch, uint32_t pitch4, int T, int* __restrict__ bad)
{
    const uint32_t wpr = pitch4 >> 2;   // words per destination row
    const long long n = (long long)rows * wpr;
    const uint32_t kadd = (uint32_t)(128 - T) * 0x01010101u;
    bool any_bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / wpr;
        const long long wd = i - r * wpr;
        const long long f0 = wd * 8;
        uint32_t v = 0;
        const uint8_t* p = src + r * src_pitch + f0;
        if (ALIGNED && f0 + 8 <= B) {
            const uint2 x = *reinterpret_cast<const uint2*>(p);
            // a byte >= T has bit 7 set either by itself or after adding 128-T (a carry out of a byte only
            // happens when that byte was >= 128 already, i.e. the word is reported anyway)
            const bool ok = (((x.x | (x.x + kadd)) | (x.y | (x.y + kadd))) & 0x80808080u) == 0u;
            if (ok) {
                // bytes b0..b3 of x.x -> nibbles 0..3, x.y -> nibbles 4..7
                const uint32_t lo = x.x & 0x0f0f0f0fu, hi = x.y & 0x0f0f0f0fu;
                const uint32_t l2 = (lo | (lo >> 4)) & 0x00ff00ffu, h2 = (hi | (hi >> 4)) & 0x00ff00ffu;
                v = ((l2 | (l2 >> 8)) & 0xffffu) | (((h2 | (h2 >> 8)) & 0xffffu) << 16);
            } else {
                any_bad = true;
#pragma unroll
                for (int f = 0; f < 8; ++f) v |= (uint32_t)min((int)p[f], T - 1) << (4 * f);
            }
        } else {
#pragma unroll
            for (int f = 0; f < 8; ++f)
                if (f0 + f < B) {
                    const int x = p[f];
                    any_bad |= x >= T;
                    v |= (uint32_t)min(x, T - 1) << (4 * f);
                }
        }
        reinterpret_cast<uint32_t*>(dst)[i] = v;
    }
    if (any_bad) atomicOr(bad, 1);
}

// Sanitised copy of a padded uint8 cluster buffer (n 16-byte vectors) for the kernel families that use the channel
// values as table indices without repacking (uint8 family, generic path): values >= T are clamped to T-1 and
// reported through *bad.
static __global__ void clamp_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long n, int T,
                                       int* __restrict__ bad)
{
    bool any_bad = false;
    const uint32_t kadd = (uint32_t)(128 - (T > 128 ? 128 : T)) * 0x01010101u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint4 v = reinterpret_cast<const uint4*>(src)[i];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        bool ok = true;
        if (T <= 128) {
#pragma unroll
            for (int q = 0; q < 4; ++q) ok &= ((w[q] | (w[q] + kadd)) & 0x80808080u) == 0u;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) ok &= (int)((w[q] >> (8 * k)) & 0xffu) < T;
        }
        if (!ok) {
            any_bad = true;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t r = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) r |= (uint32_t)min((int)((w[q] >> (8 * k)) & 0xffu), T - 1) << (8 * k);
                w[q] = r;
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4*>(dst)[i] = v;
    }
    if (any_bad) atomicOr(bad, 1);
}

// packed nibbles (rows, pitch4) -> uint8 (rows, dst_pitch): the inverse of pack_n4_kernel, for the kernel families
// that read uint8 when the caller hands in packed host buffers.
static __global__ void unpack_n4_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int rows, long long B,
                                 long long dst_pitch, uint32_t pitch4)
{
    const long long n = (long long)rows * dst_pitch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / dst_pitch, f = i - r * dst_pitch;
        dst[i] = f < B ? (uint8_t)((src[r * pitch4 + (f >> 1)] >> (4 * (f & 1))) & 15u) : (uint8_t)0;
    }
}

// Hard decisions of the first `rows` rows of a decided-cluster buffer, bit-packed: bit (f & 7) of byte f >> 3 of
// row r = (out[r][f] < threshold), i.e. the decoded bit of return_errors_all_zero (discrete_LDPC_decoder.py:297-300)
// and of the _enc drivers' comparison (WLAN/BER_simulation_OpenCL_enc.py:134).  One thread per 32 frames.
static __global__ void harddecision_bits_kernel(const uint8_t* __restrict__ out, int rows, long long B, long long pitch,
                                         int threshold, uint32_t* __restrict__ bits, long long words_per_row)
{
    const long long n = (long long)rows * words_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / words_per_row, wd = i - r * words_per_row;
        const long long f0 = wd * 32;
        const uint8_t* p = out + r * pitch + f0;
        uint32_t v = 0;
        if (f0 + 32 <= B && (pitch & 15) == 0) {
            const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 16);
            const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int k = 0; k < 4; ++k) v |= (((w[q] >> (8 * k)) & 0xffu) < (uint32_t)threshold ? 1u : 0u) << (4 * q + k);
        } else {
            for (int f = 0; f < 32; ++f)
                if (f0 + f < B) v |= (p[f] < threshold ? 1u : 0u) << f;
        }
        bits[i] = v;
    }
}

}  // namespace ibldpc
