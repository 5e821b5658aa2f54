// ib_kernels.cuh -- sm_100a kernels of the information-bottleneck (LUT) decoders.
//
// Replaces the six OpenCL kernels of Discrete_LDPC_decoding/kernels_template.cl and
// kernels_template_irreg.cl (reference file:line cited at each kernel).
//
// Data layout in HBM (all uint8, frame index fastest, `pitch` = B rounded up to 16):
//   ch  [n_var ][pitch]  channel cluster indices
//   msg [n_edge][pitch]  ONE in-place, check-node-major message array: row sc[c]+k holds the
//                        VN->CN message of the k-th neighbour of check c before a CN phase and
//                        the CN->VN message after it (the reference keeps two int32 inboxes,
//                        discrete_LDPC_decoder.py:171-173).  A variable node reaches its rows
//                        through tv[] (target_memory_cells_varnodes).
//   out [n_var ][pitch]  decided cluster indices
//
// Thread mapping: one warp = one (node, 512-frame tile); every lane moves 16 frames per
// message with one 128-bit load/store, so a warp touches 512 contiguous bytes per row.
//
// Look-ups (fast path, |T| <= 16): the iteration's stage tables are expanded in shared
// memory to one 128*W-byte row per (message m, running value t) pair, row = m*T + t, each
// lane owning one 32-bit bank column; byte (col & 3) of word (col >> 2) of that column
// holds stage `col`.  A look-up is `IMAD addr = t*RS + (m*T*RS + lane*4)` + `LDS.U8`,
// bank-conflict free for arbitrary data.  Leave-one-out chains share their common prefix
// (the sequential order of the reference is preserved, so results are bit-identical).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ibldpc {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMaxFastDc = 10;   // check-node degrees instantiated in the fast path
constexpr int kMaxFastDv = 12;   // variable-node degrees instantiated in the fast path
constexpr int kMaxGenericDeg = 64;
constexpr int kPairSlots = 16;   // lane slots of the tail-pair table (16: conflict-free LDS.64; 8: half the shared memory, 2-way conflicts)

struct IbArgs {
    // graph
    const int* __restrict__ sc;
    const int* __restrict__ deg_c;
    const int* __restrict__ sv;
    const int* __restrict__ deg_v;
    const int* __restrict__ tv;     // VN-major slot -> CN-major row
    const int* __restrict__ vidx;   // CN-major row -> variable index
    int n_var, n_chk;
    // buffers
    const uint8_t* __restrict__ ch;
    uint8_t* msg;
    uint8_t* out;
    uint32_t pitch;    // bytes per row (multiple of 16)
    uint32_t out_pitch; // packed-nibble path: bytes per row of `out` (uint8 per frame) while `pitch` is the packed pitch
    int B;             // valid frames
    int tiles;         // ceil(pitch / 512): 512-frame tiles per row
    int tpc_log2;      // log2(tiles handled by one CTA) in 0..3; a CTA's 8 warps cover 2^tpc_log2
                       // consecutive tiles of 8 >> tpc_log2 nodes; blockIdx.y selects the tile group
    // tables of this launch
    const uint8_t* __restrict__ lut;    // [nst][T*T] stage tables of this iteration (compact, reference order t*T+m)
    const uint8_t* __restrict__ match;  // [dmax][T] matching rows of this iteration or nullptr
    int T, Tc, nst, dmax_match, W, nrows, tshift;
    // iteration control
    int* flags;        // flags[it] != 0  <=> some frame had a non-zero syndrome in pass `it`
    int* inum;         // device copy of the reference's i_num
    int it;            // pass index (-1 for the iteration-0 check-node kernel)
    int early;
    int imax;
    int iter0;         // CN kernel only: read the channel values through vidx (send + iter0 fused)
    // generic path only
    const uint8_t* __restrict__ lut_all;   // whole table, reference layout
    const uint8_t* __restrict__ match_all;
    int DC, DV;
    long long vn_it_stride;  // fast path, output kernel: bytes between two iterations' VN tables
    // check-node "tail pair" table (see cn_word_pair): compact [T*T rows (a*T+b)][8 bytes] of this
    // launch, and the byte offset of its expanded copy inside the dynamic shared memory
    const uint8_t* __restrict__ pair;
    uint32_t pair_off;
    int xp_col;        // stage column stored shift-ready for the tail-pair row ("xp" encoding in the uint8 family,
                       // 4*x in the packed-nibble family), -1 = none
};

// ------------------------------------------------------------------------------------------
// shared-memory table staging
// ------------------------------------------------------------------------------------------
// "xp" encoding of a value x in [0,16): bits 2..4 = (x & 7), bit 7 = (x >= 8).  Loaded with a
// sign-extending LDS.S8 it is negative iff x >= 8, and its low five bits are the nibble shift
// 4*(x & 7), so selecting nibble x of a 64-bit row is ISETP + SEL + SHF (wrap) + LOP.
__host__ __device__ __forceinline__ uint32_t xp_encode(uint32_t x) { return ((x & 7u) << 2) | ((x & 8u) << 4); }

// Two steps: (1) the compact tables of this iteration (nst x T^2 bytes + matching rows, a few KB)
// are copied global -> shared with coalesced loads into a scratch area behind the expanded
// table; (2) every warp expands rows from that scratch copy (broadcast LDS) into the
// lane-striped layout.  Host side: dynamic smem = nrows*W*128 + stage_scratch_bytes().
__host__ __device__ __forceinline__ int stage_scratch_bytes(int nst, int T, int dmax_match)
{
    return ((nst * T * T + dmax_match * T + 15) / 16) * 16;
}

__device__ __forceinline__ void stage_tables(uint32_t* s_tab, const IbArgs& a, const uint8_t* lut)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int TT = a.T * a.T;
    const int total = a.nrows * a.W;
    uint8_t* scratch = reinterpret_cast<uint8_t*>(s_tab) + (size_t)total * 128;
    const int n_lut = a.nst * TT;
    if (((n_lut | (int)(reinterpret_cast<uintptr_t>(lut) & 3)) & 3) == 0) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(lut);
        uint32_t* dst = reinterpret_cast<uint32_t*>(scratch);
        for (int i = threadIdx.x; i < n_lut / 4; i += kThreads) dst[i] = src[i];
    } else {
        for (int i = threadIdx.x; i < n_lut; i += kThreads) scratch[i] = lut[i];
    }
    uint8_t* smatch = scratch + n_lut;
    if (a.match != nullptr)
        for (int i = threadIdx.x; i < a.dmax_match * a.T; i += kThreads) smatch[i] = a.match[i];
    __syncthreads();
    // Message alignment (MATCH, kernels_template_irreg.cl:84-96,162-173,233-241) is folded into the
    // tables: within the kernel of degree class d the last stage (d-3 for a check, d-2 for a variable
    // node) only ever produces final edge outputs, so its entries are passed through the matching row
    // of degree d here and the kernels need no separate matching look-up.  Only degree-2 checks have
    // no stage at all and keep an explicit matching column (col == nst == 0).
    const bool fold = a.match != nullptr && a.nst >= 1;
    for (int rw = warp; rw < total; rw += kWarpsPerCta) {
        const int r = rw / a.W, w = rw - r * a.W;
        const int m = r / a.T, t = r - m * a.T;
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int col = 4 * w + q;
            uint32_t e = 0;
            if (col < a.nst) {
                if (r < TT) {
                    e = scratch[col * TT + t * a.T + m];
                    if (fold && col == a.nst - 1) e = smatch[(a.dmax_match - 1) * a.T + e];
                    if (col == a.xp_col) e = xp_encode(e);
                }
            } else if (col == a.nst && a.match != nullptr && !fold) {
                if (m < a.dmax_match) e = smatch[m * a.T + t];
            }
            v |= e << (8 * q);
        }
        s_tab[rw * 32 + lane] = v;
    }
}

#define IB_SO(l) ((uint32_t)((((l) >> 2) << 7) + ((l) & 3)))   // byte offset of stage column l

__device__ __forceinline__ uint32_t lut_ld(const uint8_t* tab, uint32_t addr) { return tab[addr]; }

// byte f of `o` replaced by the low byte of `t` (one PRMT; f is a compile-time constant)
__device__ __forceinline__ uint32_t put_byte(uint32_t o, uint32_t t, int f)
{
    return f == 0 ? t : f == 1 ? __byte_perm(o, t, 0x3240u) : f == 2 ? __byte_perm(o, t, 0x3410u)
                                                                     : __byte_perm(o, t, 0x4210u);
}

// ------------------------------------------------------------------------------------------
// check node: D inputs, D leave-one-out outputs, 4 frames (one 32-bit word per message).
// Chain of kernels_template_irreg.cl:205-245 (iterations >= 1) and :60-96 (iteration 0; same
// address arithmetic when Tc == T): t = m[0]; for l: t = C[off + l*T^2 + t*T + m[l+1]].
// MATCH is a template parameter so that the 16-frame body is one basic block.
// ------------------------------------------------------------------------------------------
template <int D, bool MATCH>
__device__ __forceinline__ void cn_word(const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab,
                                        uint32_t RS, uint32_t TRS, uint32_t lane4, uint32_t match_off)
{
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        uint32_t b[D], ms[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            b[k] = __byte_perm(w[k], 0u, 0x4440u + f);   // byte f, zero-extended (one PRMT)
            ms[k] = b[k] * TRS + lane4;
        }
        uint32_t P[D > 1 ? D : 2];
        P[1] = b[0];
#pragma unroll
        for (int j = 1; j <= D - 2; ++j) P[j + 1] = lut_ld(tab, P[j] * RS + ms[j] + IB_SO(j - 1));
#pragma unroll
        for (int wo = 0; wo < D; ++wo) {
            uint32_t t = (wo == 0) ? b[1] : P[wo];
#pragma unroll
            for (int k = (wo == 0 ? 2 : wo + 1); k < D; ++k) t = lut_ld(tab, t * RS + ms[k] + IB_SO(k - 2));
            if (MATCH) t = lut_ld(tab, t * RS + match_off);
            o[wo] = put_byte(o[wo], t, f);
        }
    }
}

// Tail-pair variant (D >= 4).  All outputs w <= D-3 end with the same two look-ups
//   out_w = L_{D-3}( L_{D-4}(x_w, m_{D-2}), m_{D-1} ),
// which for one frame is a fixed function G of x_w alone (16 values -> 16 nibbles = 64 bits).
// The host composes G for every (m_{D-2}, m_{D-1}) pair (ibldpc_set_luts); the kernel fetches
// the 64-bit row once per frame with one LDS.64 and evaluates the D-2 applications with ALU
// nibble selects.  Shared-memory wavefronts per check and frame drop from
// 2(D-2) + (D-1)(D-2)/2 to (D-3) + 2 + (D-4) + (D-4)(D-3)/2 + 2   (D=6: 18 -> 12),
// the exact sequential look-up order of the reference is untouched (G is its composition).
__device__ __forceinline__ uint32_t pair_apply_xp(uint2 g, int e)
{
    const uint32_t word = e < 0 ? g.y : g.x;
    return __funnelshift_r(word, 0u, (uint32_t)e) & 15u;     // shift amount = e & 31 = 4*(x & 7)
}
// t*RS for a value held in xp form (RS = 128*W)
__device__ __forceinline__ uint32_t xp_times_rs(int e, uint32_t W)
{
    const uint32_t u = (uint32_t)e;
    return (((u & 0x1Cu) << 5) | ((u & 0x80u) << 3)) * W;
}
__device__ __forceinline__ int lut_ld_s8(const uint8_t* tab, uint32_t addr)
{
    return (int)reinterpret_cast<const signed char*>(tab)[addr];
}

template <int D>
__device__ __forceinline__ void cn_word_pair(const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab,
                                             uint32_t RS, uint32_t TRS, uint32_t lane4, uint32_t pair_base, uint32_t TPS,
                                             uint32_t PS, uint32_t W)
{
    static_assert(D >= 4, "tail-pair variant needs at least two look-up stages");
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        uint32_t b[D], ms[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            b[k] = __byte_perm(w[k], 0u, 0x4440u + f);
            ms[k] = b[k] * TRS + lane4;
        }
        // 64-bit row G(m_{D-2}, m_{D-1})[.] of the composed tail-pair table
        const uint2 g = *reinterpret_cast<const uint2*>(tab + (b[D - 2] * TPS + b[D - 1] * PS + pair_base));
        // prefix chain; the value produced by stage D-5 (P[D-3]) comes out in xp form
        uint32_t P[D];
        int Pxp = 0;                       // P[D-3] in xp form (D >= 5)
        P[1] = b[0];
#pragma unroll
        for (int j = 1; j <= D - 3; ++j) {
            const bool in_is_xp = (D >= 5) && (j == D - 3);     // P[j] = P[D-3] was produced by the xp column
            const uint32_t trs = in_is_xp ? xp_times_rs(Pxp, W) : P[j] * RS;
            if ((D >= 5) && (j - 1 == D - 5)) {                  // this look-up reads the xp column
                Pxp = lut_ld_s8(tab, trs + ms[j] + IB_SO(j - 1));
                P[j + 1] = 0;
            } else {
                P[j + 1] = lut_ld(tab, trs + ms[j] + IB_SO(j - 1));
            }
        }
        // the two outputs that do not pass through both tail stages: P[D-2] is a plain value
        // (stage D-4) unless D == 4, where P[D-2] = P[2] ... handled by the same rule below
        const uint32_t pd2_rs = P[D - 2] * RS;
        o[D - 1] = put_byte(o[D - 1], lut_ld(tab, pd2_rs + ms[D - 2] + IB_SO(D - 3)), f);
        o[D - 2] = put_byte(o[D - 2], lut_ld(tab, pd2_rs + ms[D - 1] + IB_SO(D - 3)), f);
#pragma unroll
        for (int wo = 0; wo <= D - 3; ++wo) {
            // chain up to (excluding) the two tail stages; its last look-up (stage D-5) yields xp form
            int e;
            if (wo == D - 3 && D >= 5) {
                e = Pxp;
            } else {
                uint32_t t = (wo == 0) ? b[1] : P[wo];
                bool have_xp = false;
                e = 0;
#pragma unroll
                for (int k = (wo == 0 ? 2 : wo + 1); k <= D - 3; ++k) {
                    if (k - 2 == D - 5) { e = lut_ld_s8(tab, t * RS + ms[k] + IB_SO(k - 2)); have_xp = true; }
                    else t = lut_ld(tab, t * RS + ms[k] + IB_SO(k - 2));
                }
                if (!have_xp) e = (int)(signed char)xp_encode(t);   // D == 4 (raw messages) only
            }
            o[wo] = put_byte(o[wo], pair_apply_xp(g, e), f);
        }
    }
}

template <int D, bool MATCH, bool EARLY, bool PAIR>
__device__ __forceinline__ uint32_t cn_node(const IbArgs& a, const uint8_t* tab, int s, uint32_t col,
                                            uint32_t lane4, uint32_t RS, uint32_t TRS, uint32_t valid_frames)
{
    uint4 m[D];
    if (a.iter0) {
#pragma unroll
        for (int k = 0; k < D; ++k)
            m[k] = *reinterpret_cast<const uint4*>(a.ch + (uint64_t)(uint32_t)a.vidx[s + k] * a.pitch + col);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k)
            m[k] = *reinterpret_cast<const uint4*>(a.msg + (uint64_t)(uint32_t)(s + k) * a.pitch + col);
    }
    const uint32_t match_off = (uint32_t)(D - 1) * TRS + lane4 + IB_SO(a.nst);
    uint32_t syn = 0;
    uint4 r[D];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t w[D], o[D];
#pragma unroll
        for (int k = 0; k < D; ++k) w[k] = j == 0 ? m[k].x : j == 1 ? m[k].y : j == 2 ? m[k].z : m[k].w;
        if (EARLY && !a.iter0) {
            // calc_syndrome (kernels_template_irreg.cl:304-325) on the VN->CN messages just read:
            // parity of (msg < T/2) over the D inputs, per frame byte.
            uint32_t par = 0;
            if (a.tshift >= 0) {   // T power of two: (m < T/2) == !bit(log2(T)-1)
                uint32_t x = 0;
#pragma unroll
                for (int k = 0; k < D; ++k) x ^= w[k];
                par = ((x >> a.tshift) & 0x01010101u) ^ ((D & 1) ? 0x01010101u : 0u);
            } else {
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    uint32_t p1 = 0;
#pragma unroll
                    for (int k = 0; k < D; ++k) p1 ^= (((w[k] >> (8 * f)) & 0xffu) < (uint32_t)(a.T / 2)) ? 1u : 0u;
                    par |= p1 << (8 * f);
                }
            }
            // ignore padding frames
            const int nv = (int)valid_frames - 4 * j;
            const uint32_t vmask = nv >= 4 ? 0xffffffffu : nv <= 0 ? 0u : ((1u << (8 * nv)) - 1u);
            syn |= par & vmask;
        }
        if constexpr (PAIR) cn_word_pair<D>(w, o, tab, RS, TRS, lane4, a.pair_off + (lane4 & (4u * (kPairSlots - 1))) * 2u,
                                            8u * kPairSlots * a.T, 8u * kPairSlots, (uint32_t)a.W);
        else cn_word<D, MATCH>(w, o, tab, RS, TRS, lane4, match_off);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (j == 0) r[k].x = o[k]; else if (j == 1) r[k].y = o[k]; else if (j == 2) r[k].z = o[k]; else r[k].w = o[k];
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k)
        *reinterpret_cast<uint4*>(a.msg + (uint64_t)(uint32_t)(s + k) * a.pitch + col) = r[k];
    return syn;
}

// send_channel_values_to_checknode_inbox + checknode_update_iter0 (kernels_template_irreg.cl:13-99)
// when a.iter0, checknode_update + calc_syndrome (:181-246, :304-325) otherwise.
// One instantiation per check-node degree; `nodes` lists the checks of that degree, so every
// launch has exactly the register budget its degree needs.
template <int D, bool MATCH, bool EARLY, bool PAIR>
__global__ void __launch_bounds__(kThreads, (D <= 6 ? (PAIR ? 3 : 4) : (D <= 8 ? 3 : 2)))
ib_cn_fast_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    if (EARLY && a.it >= 1 && a.flags[a.it - 1] == 0) return;   // batch already converged
    stage_tables(s_tab, a, a.lut);
    if (PAIR) {
        // expand the composed tail-pair rows: 16 lane slots x 8 bytes per (a,b) row, so that the two
        // half-warp phases of an LDS.64 are bank-conflict free for arbitrary data
        const uint2* src = reinterpret_cast<const uint2*>(a.pair);
        uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(s_tab) + a.pair_off);
        const int n = a.T * a.T * kPairSlots;
        for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = src[i / kPairSlots];
    }
    __syncthreads();
    const uint8_t* tab = reinterpret_cast<const uint8_t*>(s_tab);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane4 = lane * 4, RS = 128u * a.W, TRS = RS * a.T;
    // CTA -> (tile group, node subset): no divisions, 32-bit offsets
    const int tile = (blockIdx.y << a.tpc_log2) + (warp & ((1 << a.tpc_log2) - 1));
    const int nps = kWarpsPerCta >> a.tpc_log2;          // nodes per CTA step
    const int stride = gridDim.x * nps;
    const uint32_t col = ((uint32_t)tile * 32u + lane) * 16u;
    uint32_t syn = 0;
    if (tile < a.tiles && col < a.pitch) {
        const int vf = a.B - (int)col;
        const uint32_t valid = vf >= 16 ? 16u : vf <= 0 ? 0u : (uint32_t)vf;
        int i = blockIdx.x * nps + (warp >> a.tpc_log2);
        int s = i < n_nodes ? a.sc[nodes[i]] : 0;
        while (i < n_nodes) {
            // the next node's slot index is fetched while this node is being computed
            const int i2 = i + stride;
            const int s2 = i2 < n_nodes ? a.sc[nodes[i2]] : 0;
            syn |= cn_node<D, MATCH, EARLY, PAIR>(a, tab, s, col, lane4, RS, TRS, valid);
            i = i2;
            s = s2;
        }
    }
    if (EARLY && !a.iter0) {
        // warp-ballot syndrome check: one flag write per warp that saw an unsatisfied check
        const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
        if (any != 0 && lane == 0) atomicOr(&a.flags[a.it], 1);
    }
}

// ------------------------------------------------------------------------------------------
// variable node: channel value + D inbox messages.
// Update chain kernels_template_irreg.cl:125-177, decision chain :277-300:
//   t = V[off + m0*T + m1]; for l>=1: t = V[off + Tc*T + (l-1)*T^2 + t*T + m[l+1]].
// ------------------------------------------------------------------------------------------
template <int D, bool MATCH, bool DECIDE>
__device__ __forceinline__ void vn_word(uint32_t chw, const uint32_t (&w)[D], uint32_t (&o)[D], uint32_t& dec,
                                        const uint8_t* tab, uint32_t RS, uint32_t TRS, uint32_t lane4,
                                        uint32_t match_off)
{
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
    dec = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        uint32_t ms[D + 1];   // ms[k] for y_k, k = 1..D
#pragma unroll
        for (int k = 1; k <= D; ++k) ms[k] = __byte_perm(w[k - 1], 0u, 0x4440u + f) * TRS + lane4;
        uint32_t P[D + 2];
        P[1] = __byte_perm(chw, 0u, 0x4440u + f);
#pragma unroll
        for (int j = 1; j <= D - 1; ++j) P[j + 1] = lut_ld(tab, P[j] * RS + ms[j] + IB_SO(j - 1));
        if (DECIDE) {
            dec = put_byte(dec, lut_ld(tab, P[D] * RS + ms[D] + IB_SO(D - 1)), f);
        } else {
#pragma unroll
            for (int wo = 1; wo <= D; ++wo) {
                uint32_t t = P[wo];
#pragma unroll
                for (int k = wo + 1; k <= D; ++k) t = lut_ld(tab, t * RS + ms[k] + IB_SO(k - 2));
                if (MATCH) t = lut_ld(tab, t * RS + match_off);
                o[wo - 1] = put_byte(o[wo - 1], t, f);
            }
        }
    }
}

template <int D> struct VnIn { uint4 c4; uint4 m[D]; };
template <int D> struct VnIdx { int v; int rows[D]; bool ok; };

template <int D>
__device__ __forceinline__ void vn_load_idx(const IbArgs& a, const int* __restrict__ nodes, int n_nodes, int i, VnIdx<D>& x)
{
    x.ok = i < n_nodes;
    x.v = 0;
#pragma unroll
    for (int k = 0; k < D; ++k) x.rows[k] = 0;
    if (x.ok) {
        x.v = nodes[i];
        const int s = a.sv[x.v];
#pragma unroll
        for (int k = 0; k < D; ++k) x.rows[k] = a.tv[s + k];
    }
}

template <int D>
__device__ __forceinline__ void vn_load_msgs(const IbArgs& a, const VnIdx<D>& x, uint32_t col, VnIn<D>& in)
{
    if (x.ok) {
        in.c4 = *reinterpret_cast<const uint4*>(a.ch + (uint64_t)(uint32_t)x.v * a.pitch + col);
#pragma unroll
        for (int k = 0; k < D; ++k)
            in.m[k] = *reinterpret_cast<const uint4*>(a.msg + (uint64_t)(uint32_t)x.rows[k] * a.pitch + col);
    }
}

template <int D, bool MATCH, bool DECIDE>
__device__ __forceinline__ void vn_compute_store(const IbArgs& a, const uint8_t* tab, const VnIdx<D>& x, const VnIn<D>& in,
                                                 uint32_t col, uint32_t lane4, uint32_t RS, uint32_t TRS)
{
    if (!DECIDE && D == 1) {   // degree-1 variable node forwards the raw channel value (:132-136)
        *reinterpret_cast<uint4*>(a.msg + (uint64_t)(uint32_t)x.rows[0] * a.pitch + col) = in.c4;
        return;
    }
    const uint32_t match_off = (uint32_t)(D - 1) * TRS + lane4 + IB_SO(a.nst);
    uint4 r[D];
    uint4 dec4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t w[D], o[D], dec;
        const uint32_t chw = j == 0 ? in.c4.x : j == 1 ? in.c4.y : j == 2 ? in.c4.z : in.c4.w;
#pragma unroll
        for (int k = 0; k < D; ++k) w[k] = j == 0 ? in.m[k].x : j == 1 ? in.m[k].y : j == 2 ? in.m[k].z : in.m[k].w;
        vn_word<D, MATCH, DECIDE>(chw, w, o, dec, tab, RS, TRS, lane4, match_off);
        if (DECIDE) {
            if (j == 0) dec4.x = dec; else if (j == 1) dec4.y = dec; else if (j == 2) dec4.z = dec; else dec4.w = dec;
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                if (j == 0) r[k].x = o[k]; else if (j == 1) r[k].y = o[k]; else if (j == 2) r[k].z = o[k]; else r[k].w = o[k];
            }
        }
    }
    if (DECIDE) {
        *reinterpret_cast<uint4*>(a.out + (uint64_t)(uint32_t)x.v * a.pitch + col) = dec4;
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) *reinterpret_cast<uint4*>(a.msg + (uint64_t)(uint32_t)x.rows[k] * a.pitch + col) = r[k];
    }
}

// Software pipeline over the node list of one degree class.  The variable-node phase is a
// gather (row indices through sv[] and tv[]: three dependent loads before the first message
// byte arrives), so both the indices (two nodes ahead) and -- for small degrees -- the messages
// (one node ahead, ping-pong register buffers) are fetched while the current node is computed.
// Every message row belongs to exactly one variable node, so prefetching the next node's rows
// before the current node's results are stored cannot alias.
template <int D, bool MATCH, bool DECIDE>
__device__ __forceinline__ void vn_loop(const IbArgs& a, const uint8_t* tab, const int* __restrict__ nodes, int n_nodes)
{
    // Measured on B200, (3,6) n=8000, B=16384: message double-buffering costs 16 registers per
    // thread (80 instead of 64 -> 3 instead of 4 CTAs per SM) and is slower (0.183 ms vs 0.172 ms per
    // launch) than index prefetch alone, so it is compiled out.
    constexpr bool kPrefetchMsgs = false;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lane4 = lane * 4, RS = 128u * a.W, TRS = RS * a.T;
    const int tile = (blockIdx.y << a.tpc_log2) + (warp & ((1 << a.tpc_log2) - 1));
    const int nps = kWarpsPerCta >> a.tpc_log2;
    const int stride = gridDim.x * nps;
    const uint32_t col = ((uint32_t)tile * 32u + lane) * 16u;
    if (tile >= a.tiles || col >= a.pitch) return;
    int i = blockIdx.x * nps + (warp >> a.tpc_log2);
    VnIdx<D> cur, nxt, nn;
    vn_load_idx<D>(a, nodes, n_nodes, i, cur);
    vn_load_idx<D>(a, nodes, n_nodes, i + stride, nxt);
    int inn = i + 2 * stride;
    if (kPrefetchMsgs) {
        VnIn<D> bufA, bufB;
        vn_load_msgs<D>(a, cur, col, bufA);
        while (cur.ok) {
            vn_load_msgs<D>(a, nxt, col, bufB);
            vn_load_idx<D>(a, nodes, n_nodes, inn, nn);
            inn += stride;
            vn_compute_store<D, MATCH, DECIDE>(a, tab, cur, bufA, col, lane4, RS, TRS);
            cur = nxt;
            nxt = nn;
            if (!cur.ok) break;
            vn_load_msgs<D>(a, nxt, col, bufA);
            vn_load_idx<D>(a, nodes, n_nodes, inn, nn);
            inn += stride;
            vn_compute_store<D, MATCH, DECIDE>(a, tab, cur, bufB, col, lane4, RS, TRS);
            cur = nxt;
            nxt = nn;
        }
    } else {
        while (cur.ok) {
            VnIn<D> buf;
            vn_load_msgs<D>(a, cur, col, buf);
            vn_load_idx<D>(a, nodes, n_nodes, inn, nn);
            inn += stride;
            vn_compute_store<D, MATCH, DECIDE>(a, tab, cur, buf, col, lane4, RS, TRS);
            cur = nxt;
            nxt = nn;
        }
    }
}

// Number of executed passes under the reference's stop rule (discrete_LDPC_decoder.py:233-276):
// pass `it` runs iff it == 0 or pass it-1 left a non-zero syndrome somewhere in the batch.
__device__ __forceinline__ int executed_passes(const IbArgs& a)
{
    int passes = a.imax - 1;
    if (a.early)
        for (int it = 0; it < a.imax - 1; ++it)
            if (a.flags[it] == 0) { passes = it + 1; break; }
    return passes;
}

// varnode_update (kernels_template_irreg.cl:103-179); one instantiation per variable-node degree.
template <int D, bool MATCH>
__global__ void __launch_bounds__(kThreads, (D <= 4 ? 4 : (D <= 8 ? 3 : 2)))   // 5 CTAs/SM (<= 51 regs) spills and is 1.5x slower
ib_vn_fast_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;
    if (D > 1) {   // degree-1 nodes forward the channel value, no tables needed
        stage_tables(s_tab, a, a.lut);
        __syncthreads();
    }
    vn_loop<D, MATCH, false>(a, reinterpret_cast<const uint8_t*>(s_tab), nodes, n_nodes);
}

// calc_varnode_output (kernels_template_irreg.cl:249-302) with the VN table of iteration i_num-1
template <int D>
__global__ void __launch_bounds__(kThreads) ib_out_fast_kernel(IbArgs a, const int* __restrict__ nodes, int n_nodes)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    __shared__ int s_passes;
    if (threadIdx.x == 0) {
        s_passes = executed_passes(a);
        if (blockIdx.x == 0 && blockIdx.y == 0) *a.inum = s_passes + 1;
    }
    __syncthreads();
    stage_tables(s_tab, a, a.lut + (long long)s_passes * a.vn_it_stride);
    __syncthreads();
    vn_loop<D, false, true>(a, reinterpret_cast<const uint8_t*>(s_tab), nodes, n_nodes);
}

// ------------------------------------------------------------------------------------------
// generic path: any |T| <= 256, Tc != T, degrees <= 64.  One thread per (node, frame), tables
// read from global memory with the reference's index arithmetic.  Same in-place layout.
// ------------------------------------------------------------------------------------------
static __global__ void ib_cn_generic_kernel(IbArgs a)
{
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int T = a.T, Tc = a.Tc;
    const uint8_t* C = a.lut_all;
    bool bad = false;
    if (f < a.pitch) {
        for (int c = blockIdx.y; c < a.n_chk; c += gridDim.y) {
            const int d = a.deg_c[c], s = a.sc[c];
            uint8_t in[kMaxGenericDeg];
            int par = 0;
            for (int k = 0; k < d; ++k) {
                in[k] = a.iter0 ? a.ch[(long long)a.vidx[s + k] * a.pitch + f] : a.msg[(long long)(s + k) * a.pitch + f];
                par ^= (in[k] < T / 2);
            }
            if (par && f < a.B) bad = true;
            for (int w = 0; w < d; ++w) {
                int t;
                if (a.iter0) {   // kernels_template_irreg.cl:60-96
                    int i0 = (w == 0) ? 1 : 0, i1 = (w <= 1) ? 2 : 1;
                    t = (d >= 3) ? C[in[i0] * Tc + in[i1]] : in[i0];
                    int l = 1;
                    for (int k = i1 + 1; k < d; ++k) {
                        if (k == w) continue;
                        t = C[t * T + in[k] + Tc * Tc + (l - 1) * Tc * T];
                        ++l;
                    }
                    if (a.match_all) t = a.match_all[(d - 1) * T + t];
                } else {         // :205-245
                    const int off = Tc * Tc + (a.DC - 3) * Tc * T + a.it * ((a.DC - 2) * T * T);
                    int first = (w == 0) ? 1 : 0;
                    t = in[first];
                    int l = 0;
                    for (int k = first + 1; k < d; ++k) {
                        if (k == w) continue;
                        t = C[off + t * T + in[k] + l * T * T];
                        ++l;
                    }
                    if (a.match_all) t = a.match_all[(a.it + 1) * T * a.DC + (d - 1) * T + t];
                }
                a.msg[(long long)(s + w) * a.pitch + f] = (uint8_t)t;
            }
        }
    }
    if (a.early && !a.iter0) {
        const unsigned any = __ballot_sync(0xffffffffu, bad);
        if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[a.it], 1);
    }
}

template <bool DECIDE>
__global__ void ib_vn_generic_kernel(IbArgs a)
{
    __shared__ int s_passes;
    int it = a.it;
    if (DECIDE) {
        if (threadIdx.x == 0) {
            s_passes = executed_passes(a);
            if (blockIdx.x == 0 && blockIdx.y == 0) *a.inum = s_passes + 1;
        }
        __syncthreads();
        it = s_passes;
    } else if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) {
        return;
    }
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.pitch) return;
    const int T = a.T, Tc = a.Tc;
    const uint8_t* V = a.lut_all;
    const int off = it * (Tc * T + (a.DV - 1) * T * T);
    for (int v = blockIdx.y; v < a.n_var; v += gridDim.y) {
        const int d = a.deg_v[v], s = a.sv[v];
        const int m0 = a.ch[(long long)v * a.pitch + f];
        uint8_t in[kMaxGenericDeg];
        int rows[kMaxGenericDeg];
        for (int k = 0; k < d; ++k) {
            rows[k] = a.tv[s + k];
            in[k] = a.msg[(long long)rows[k] * a.pitch + f];
        }
        if (DECIDE) {            // :277-300
            int t = V[off + m0 * T + in[0]];
            for (int l = 1; l < d; ++l) t = V[off + t * T + in[l] + Tc * T + (l - 1) * T * T];
            a.out[(long long)v * a.pitch + f] = (uint8_t)t;
        } else if (d == 1) {     // :132-136
            a.msg[(long long)rows[0] * a.pitch + f] = (uint8_t)m0;
        } else {                 // :137-177
            for (int w = 0; w < d; ++w) {
                int first = (w == 0) ? 1 : 0;
                int t = V[off + m0 * T + in[first]];
                int l = 1;
                for (int k = first + 1; k < d; ++k) {
                    if (k == w) continue;
                    t = V[off + t * T + in[k] + Tc * T + (l - 1) * T * T];
                    ++l;
                }
                if (a.match_all) t = a.match_all[it * T * a.DV + (d - 1) * T + t];
                a.msg[(long long)rows[w] * a.pitch + f] = (uint8_t)t;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// pitch helpers: (rows, B) contiguous <-> (rows, pitch) padded
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pad_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, long long rows, long long B,
                                long long pitch, T fill)
{
    const long long n = rows * pitch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / pitch, c = i - r * pitch;
        dst[i] = c < B ? src[r * B + c] : fill;
    }
}
template <typename T>
__global__ void unpad_rows_kernel(const T* __restrict__ src, T* __restrict__ dst, long long rows, long long B,
                                  long long pitch)
{
    const long long n = rows * B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / B, c = i - r * B;
        dst[i] = src[r * pitch + c];
    }
}

}  // namespace ibldpc
