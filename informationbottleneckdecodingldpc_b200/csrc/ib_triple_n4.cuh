// ib_triple_n4.cuh -- degree-3 variable nodes through ONE three-input table (packed-nibble family).
//
// A variable node of degree 3 sends  out_1 = S1(S0(ch, y2), y3),  out_2 = S1(S0(ch, y1), y3),  out_3 = S1(S0(ch, y1), y2)
// (kernels_template_irreg.cl:125-177 with the matching row of degree 3 folded into S1): three evaluations of the SAME
// function F(a, b, c) = S1(S0(a, b), c) of three four-bit arguments.  F has 16^3 = 4096 one-byte entries; replicated per
// lane it is 128 KB of shared memory (bank = lane: conflict-free for arbitrary data), which a 227 KB SM holds next to
// nothing else -- and the update costs THREE look-ups per frame instead of five.  The (3,6) kernels are bound by the
// shared-memory look-up pipe (92 % busy), so the wavefronts are what counts.
// F is composed on the host for every iteration (ibldpc_set_luts: h_vn3); values and results are bit-identical.
//
// Layout: the channel value a is the MINOR index -- it is the one argument all three look-ups of a frame share, so its
// (non-linear) word / byte split is computed once per frame: entry i = b*256 + c*16 + a lives in byte (a & 3) of word
// [i >> 2][lane]:
//   addr = (b*16 + c) * 512 + (a >> 2) * 128 + lane * 4 + (a & 3).
#pragma once
#include "ib_kernels_n4.cuh"

namespace ibldpc {

constexpr int kTripleEntries = kTS * kTS * kTS;          // 4096
constexpr int kTripleBytes = kTripleEntries * 32;        // 128 KB

// the (c >> 2) and (c & 3) bytes of the eight nibbles of a word, even / odd frames
struct SplitBytes { NibBytes hi, lo; };
__device__ __forceinline__ SplitBytes split_bytes(uint32_t w)
{
    SplitBytes s;
    s.hi.e = (w >> 2) & 0x03030303u;
    s.hi.o = (w >> 6) & 0x03030303u;
    s.lo.e = w & 0x03030303u;
    s.lo.o = (w >> 4) & 0x03030303u;
    return s;
}

__device__ __forceinline__ void vn3_word_n4(uint32_t chw, const uint32_t (&w)[3], uint32_t (&o)[3], const uint8_t* tab, uint32_t lane4)
{
    const NibBytes p23 = pair_bytes(w[1], w[2]), p13 = pair_bytes(w[0], w[2]), p12 = pair_bytes(w[0], w[1]);   // (b << 4 | c)
    const SplitBytes ca = split_bytes(chw);
    o[0] = o[1] = o[2] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        const uint32_t ua = byte_mad<128u>(ca.hi, f, byte_mad<1u>(ca.lo, f, lane4));   // channel term, once per frame
        o[0] += (uint32_t)tab[byte_mad<128u>(p23, f, 0u) * 4u + ua] << (4 * f);   // edge 1: ch, y2, y3
        o[1] += (uint32_t)tab[byte_mad<128u>(p13, f, 0u) * 4u + ua] << (4 * f);   // edge 2: ch, y1, y3
        o[2] += (uint32_t)tab[byte_mad<128u>(p12, f, 0u) * 4u + ua] << (4 * f);   // edge 3: ch, y1, y2
    }
}

// ---- degree-6 check nodes ------------------------------------------------------------------------------------
// The first two stages of every leave-one-out chain of a degree-6 check are one evaluation of
//   F(x, y, z) = 4 * C1( C0(x, y), z )                      (kernels_template_irreg.cl:205-245, stages 0 and 1)
// with (x, y, z) = (m1,m2,m3), (m0,m2,m3), (m0,m1,m3), (m0,m1,m2) for the outputs 0..3 (the last one is also the prefix
// P3 of the chain).  With the tail-pair row of the last two stages that leaves 7 one-byte look-ups and one LDS.64 per
// frame instead of 10 + 1.  Layout: x minor (two of the four look-ups share x = m0):
//   entry (y*16 + z)*16 + x  ->  addr = (y*16 + z) * 512 + (x >> 2) * 128 + lane * 4 + (x & 3);  values are stored as 4 * v
// (the nibble shift into the tail-pair row; as a table index they are scaled by TRS / 4).
// Shared memory of the kernel: [tail-pair rows 32 KB][F 128 KB][stage tables 32 KB + staging scratch].
__device__ __forceinline__ void cn6_word_n4_triple(const uint32_t (&w)[6], uint32_t (&o)[6], const uint8_t* tab, const uint8_t* ptab,
                                                   const uint8_t* ttab, uint32_t lane4, uint32_t slot8)
{
    constexpr uint32_t RS = 128u, TRS = RS * kTS, PS = 8u * kPairSlots;
    const NibBytes p23 = pair_bytes(w[2], w[3]), p13 = pair_bytes(w[1], w[3]), p12 = pair_bytes(w[1], w[2]);
    const SplitBytes x0 = split_bytes(w[0]), x1 = split_bytes(w[1]);
    const NibBytes b3 = nib_bytes<1>(w[3]), b4 = nib_bytes<1>(w[4]), b5 = nib_bytes<1>(w[5]);
    const NibBytes pb = pair_bytes(w[4], w[5]);
#pragma unroll
    for (int k = 0; k < 6; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        const uint32_t u0 = byte_mad<128u>(x0.hi, f, byte_mad<1u>(x0.lo, f, lane4));
        const uint32_t u1 = byte_mad<128u>(x1.hi, f, byte_mad<1u>(x1.lo, f, lane4));
        const uint32_t t23 = byte_mad<128u>(p23, f, 0u) * 4u, t13 = byte_mad<128u>(p13, f, 0u) * 4u, t12 = byte_mad<128u>(p12, f, 0u) * 4u;
        const uint2 g2 = *reinterpret_cast<const uint2*>(ptab + byte_mad<PS>(pb, f, slot8));
        const unsigned long long g = ((unsigned long long)g2.y << 32) | g2.x;
        const uint32_t e0 = ttab[t23 + u1];   // output 0: m1, m2, m3
        const uint32_t e1 = ttab[t23 + u0];   // output 1: m0, m2, m3
        const uint32_t e2 = ttab[t13 + u0];   // output 2: m0, m1, m3
        const uint32_t e3 = ttab[t12 + u0];   // output 3 and prefix P3: m0, m1, m2
        const uint32_t P4 = lut_ld(tab, e3 * (TRS / 4u) + byte_mad<128u>(b3, f, lane4) + IB_SO(2));
        o[5] += lut_ld(tab, P4 * TRS + byte_mad<128u>(b4, f, lane4) + IB_SO(3)) << (4 * f);
        o[4] += lut_ld(tab, P4 * TRS + byte_mad<128u>(b5, f, lane4) + IB_SO(3)) << (4 * f);
        nib_push(o[0], g, e0, f);
        nib_push(o[1], g, e1, f);
        nib_push(o[2], g, e2, f);
        nib_push(o[3], g, e3, f);
    }
}

// ---- check nodes of degree 7 and 8 ---------------------------------------------------------------------------
// Same F (it only depends on the stage tables 0 and 1, which every degree class shares): the four chains that start with
// three of (m0..m3) take their first two stages from F and continue with the stage columns 2.. (D = 7: 12 + 1 look-ups
// per frame instead of 15 + 1, D = 8: 18 + 1 instead of 21 + 1).  The kernel stages the columns 2..D-3 only -- LOCAL column
// c is stage column c + 2 -- so the table set stays one word (32 KB) next to F and the tail-pair rows.
// F values are 4 * v: as a table index they are scaled by TRS / 4, exactly like the values of the "x4" column D - 5.
template <int D>
__device__ __forceinline__ void cn_word_n4_triple(const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab, const uint8_t* ptab,
                                                  const uint8_t* ttab, uint32_t lane4, uint32_t slot8)
{
    static_assert(D == 7 || D == 8, "generic three-input-table variant: degrees 7 and 8 (degree 6: cn6_word_n4_triple)");
    constexpr uint32_t RS = 128u, TRS = RS * kTS, PS = 8u * kPairSlots;
    constexpr int XP = D - 5;                        // stage column in x4 form (feeds the tail-pair row)
    const NibBytes p23 = pair_bytes(w[2], w[3]), p13 = pair_bytes(w[1], w[3]), p12 = pair_bytes(w[1], w[2]);
    const SplitBytes x0 = split_bytes(w[0]), x1 = split_bytes(w[1]);
    NibBytes b[D];
#pragma unroll
    for (int k = 3; k < D; ++k) b[k] = nib_bytes<1>(w[k]);
    const NibBytes pb = pair_bytes(w[D - 2], w[D - 1]);
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
        const uint32_t u0 = byte_mad<128u>(x0.hi, f, byte_mad<1u>(x0.lo, f, lane4));
        const uint32_t u1 = byte_mad<128u>(x1.hi, f, byte_mad<1u>(x1.lo, f, lane4));
        const uint32_t t23 = byte_mad<128u>(p23, f, 0u) * 4u, t13 = byte_mad<128u>(p13, f, 0u) * 4u, t12 = byte_mad<128u>(p12, f, 0u) * 4u;
        uint32_t ms[D];
#pragma unroll
        for (int k = 3; k < D; ++k) ms[k] = byte_mad<128u>(b[k], f, lane4);
        const uint2 g2 = *reinterpret_cast<const uint2*>(ptab + byte_mad<PS>(pb, f, slot8));
        const unsigned long long g = ((unsigned long long)g2.y << 32) | g2.x;
        // c[wo]: state of the chain of output wo after its first three inputs, 4 * v; c[3] is also the prefix P[3]
        uint32_t c[4];
        c[0] = ttab[t23 + u1];   // m1, m2, m3
        c[1] = ttab[t23 + u0];   // m0, m2, m3
        c[2] = ttab[t13 + u0];   // m0, m1, m3
        c[3] = ttab[t12 + u0];   // m0, m1, m2
        // prefix chain P[j + 1] = S_{j-1}(P[j], m_j), j = 3 .. D-3 (stage column j - 1 = local column j - 3)
        uint32_t P[D];
        P[3] = c[3];
#pragma unroll
        for (int j = 3; j <= D - 3; ++j) {
            const bool in_x4 = (j == 3) || (j - 2 == XP);          // P[j] came out of F or out of the x4 column
            P[j + 1] = lut_ld(tab, P[j] * (in_x4 ? TRS / 4u : TRS) + ms[j] + IB_SO(j - 3));
        }
        // the two outputs that skip one of the tail messages (P[D-2] came out of column D-4: plain)
        o[D - 1] += lut_ld(tab, P[D - 2] * TRS + ms[D - 2] + IB_SO(D - 5)) << (4 * f);
        o[D - 2] += lut_ld(tab, P[D - 2] * TRS + ms[D - 1] + IB_SO(D - 5)) << (4 * f);
#pragma unroll
        for (int wo = 0; wo <= D - 3; ++wo) {
            uint32_t e;   // 4 * x_w
            if (wo == D - 3) {
                e = P[D - 3];
            } else {
                uint32_t t = wo <= 3 ? c[wo] : P[wo];
                bool x4 = wo <= 3 || (wo - 2 == XP);                 // wo >= 4: P[wo] came out of column wo - 2
#pragma unroll
                for (int k = (wo <= 3 ? 4 : wo + 1); k <= D - 3; ++k) {
                    t = lut_ld(tab, t * (x4 ? TRS / 4u : TRS) + ms[k] + IB_SO(k - 4));
                    x4 = (k - 2 == XP);
                }
                e = t;   // the last look-up read column D-5
            }
            nib_push(o[wo], g, e, f);
        }
    }
}

// checknode_update (+ iteration 0, + syndrome) of one degree class through the three-input table; a.lut_all = F of this
// table block (4096 bytes), a.pair = tail-pair rows of the class, a.lut / a.nst / a.xp_col = the stage columns the class
// still reads: degree 6 all four (only the columns 2 and 3 are read), degrees 7 and 8 the columns 2..D-3 as local 0..
template <int D, bool EARLY, int NT>
__global__ void __launch_bounds__(NT, 1) ib_cn_n4_tri_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_all[];
    if (EARLY && a.it >= 1 && a.flags[a.it - 1] == 0) return;   // batch already converged
    uint32_t* s_tri = s_all + kPairBytes / 4;
    uint32_t* s_tab = s_tri + kTripleBytes / 4;
    {
        const uint2* src = reinterpret_cast<const uint2*>(a.pair);
        uint2* dst = reinterpret_cast<uint2*>(s_all);
        for (int i = threadIdx.x; i < kTS * kTS * kPairSlots; i += NT) {
            const int r = i / kPairSlots, ra = r / kTS, rb = r - ra * kTS;
            dst[i] = (ra < a.T && rb < a.T) ? src[ra * a.T + rb] : make_uint2(0u, 0u);
        }
        const uint32_t* __restrict__ f3 = reinterpret_cast<const uint32_t*>(a.lut_all);
        for (int q = threadIdx.x; q < kTripleBytes / 4; q += NT) s_tri[q] = f3[q >> 5];
    }
    stage_tables_n4<1, NT>(s_tab, a, a.lut);
    __syncthreads();
    const uint32_t syn = cn_loop_n4<D, false, EARLY, 2, true, NT, 1>(a, reinterpret_cast<const uint8_t*>(s_tab),
                                                                     reinterpret_cast<const uint8_t*>(s_all), nodes, n_nodes);
    if (EARLY && !a.iter0) {
        const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
        if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[a.it], 1);
    }
}

// varnode_update of the degree-3 class (kernels_template_irreg.cl:103-179); a.lut = F of this iteration (4096 bytes).
// Node loop and software pipeline of vn_loop_n4; VEC words (8 VEC frames) per lane and message.
template <int VEC, int NT>
__global__ void __launch_bounds__(NT, 1) ib_vn3_n4_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;
    {
        const uint32_t* __restrict__ src = reinterpret_cast<const uint32_t*>(a.lut);
        for (int q = threadIdx.x; q < kTripleBytes / 4; q += NT) s_tab[q] = src[q >> 5];
    }
    __syncthreads();
    const uint8_t* tab = reinterpret_cast<const uint8_t*>(s_tab);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane4 = lane * 4;
    const int tile = (blockIdx.y << a.tpc_log2) + (warp & ((1 << a.tpc_log2) - 1));
    const int nps = (NT / 32) >> a.tpc_log2;
    const int stride = gridDim.x * nps;
    const uint32_t col = ((uint32_t)tile * 32u + lane) * (4u * VEC);
    if (tile >= a.tiles || col >= a.pitch) return;
    int i = blockIdx.x * nps + (warp >> a.tpc_log2);
    VnIdx<3> cur, nxt, nn;
    vn_load_idx<3>(a, nodes, n_nodes, i, cur);
    vn_load_idx<3>(a, nodes, n_nodes, i + stride, nxt);
    int inn = i + 2 * stride;
    while (cur.ok) {
        VnIn4<3, VEC> buf;
        vn_load_msgs_n4<3, VEC>(a, cur, col, buf);
        vn_load_idx<3>(a, nodes, n_nodes, inn, nn);
        inn += stride;
        uint32_t r[3][VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const uint32_t w[3] = {buf.m[0][j], buf.m[1][j], buf.m[2][j]};
            uint32_t o[3];
            vn3_word_n4(buf.c[j], w, o, tab, lane4);
#pragma unroll
            for (int k = 0; k < 3; ++k) r[k][j] = o[k];
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) st_words<VEC>(a.msg + (uint64_t)(uint32_t)cur.rows[k] * a.pitch + col, r[k]);
        cur = nxt;
        nxt = nn;
    }
}

}  // namespace ibldpc
