// ib_kernels_t32.cuh -- shared-memory kernel family for 16 < |T| <= 32 (cardinality_T_channel = cardinality_T_decoder_ops
// = 32 is the reference's own 802.11n design, Irregular_LDPC_Decoding/WLAN/decoder_config_generation.py:25-26).
//
// Same flooding schedule, same in-place check-node-major message array and the same look-up order as the |T| <= 16
// families (bit-identical to kernels_template_irreg.cl:33-99, :103-179, :181-246, :249-325); messages and channel
// values are one byte per frame (uint8 family layout of ib_kernels.cuh).
//
// Tables.  A 32 x 32 stage table is 1 KB = 8 words per bank: un-replicated, 32 lanes with arbitrary (t, m) pairs need
// ~3.5 wavefronts per look-up.  Replicated per lane it is conflict-free:
//   striped stage  (32 KB)  byte (m & 3) of word [t][m >> 2][lane]:  addr = t*1024 + (m >> 2)*128 + lane*4 + (m & 3)
//   plain stage    ( 1 KB)  byte [t][m]:                             addr = t*32 + m
// A class of degree D reaches nst = D-2 (check), D-1 (variable update) or D (decision) stages; the LAST min(nst, 6)
// stages are striped (192 KB) -- the leave-one-out chains use stage c about c+2 times, so the late stages carry most of
// the look-ups (d_v = 11: 51 of 65) -- and the early ones stay plain.  Message alignment is folded into the last
// stage.  The image of every (iteration, degree class) is expanded on the host at ibldpc_set_luts and brought in with
// one cp.async.bulk (TMA bulk copy) per CTA, like the fused per-phase kernels of the packed family.
//
// One launch per phase and degree class, persistent 1024-thread CTAs (one per SM), (node, tile) items pulled from a
// shared-memory counter.
#pragma once
#include "ib_phase_n4.cuh"   // bulk-copy helpers, ld_words / st_words

namespace ibldpc {

constexpr int kT32Threads = 1024;
constexpr int kT32MaxStriped = 6;
constexpr int kT32StripedBytes = 32 * 1024, kT32PlainBytes = 1024;

__host__ __device__ constexpr int t32_cols(int mode, int d) { return mode == kPhaseCn ? d - 2 : mode == kPhaseVn ? d - 1 : d; }
__host__ __device__ constexpr int t32_striped(int nst) { return nst < kT32MaxStriped ? nst : kT32MaxStriped; }
// image = [striped stages (the last ones)][plain stages (the first ones)]
__host__ __device__ constexpr int t32_image_bytes(int nst)
{
    const int b = t32_striped(nst) * kT32StripedBytes + (nst - t32_striped(nst)) * kT32PlainBytes;
    return b < 16 ? 16 : b;
}
__host__ __device__ constexpr bool t32_is_striped(int nst, int col) { return col >= nst - t32_striped(nst); }
__host__ __device__ constexpr int t32_col_base(int nst, int col)
{
    const int np = nst - t32_striped(nst);
    return t32_is_striped(nst, col) ? (col - np) * kT32StripedBytes : t32_striped(nst) * kT32StripedBytes + col * kT32PlainBytes;
}
// words (4 frames each) a lane moves per message row
__host__ __device__ constexpr int t32_vec(int mode, int d) { return mode == kPhaseCn ? (d <= 8 ? 2 : 1) : (d <= 4 ? 4 : d <= 5 ? 2 : 1); }

struct T32Args {
    IbArgs a;
    const uint8_t* image;       // this launch's table image in global memory
    long long image_stride;     // decision kernel: bytes between the images of consecutive iterations
    const int* nodes;
    const int* starts;
    int n_nodes;
};

// look-up in stage `COL` of a class with NST stages: t = running value, ms / mp = striped / plain message offsets
template <int NST, int COL>
__device__ __forceinline__ uint32_t t32_lut(const uint8_t* tab, uint32_t t, uint32_t ms, uint32_t mp)
{
    if constexpr (t32_is_striped(NST, COL)) return tab[t32_col_base(NST, COL) + t * 1024u + ms];
    else return tab[t32_col_base(NST, COL) + t * 32u + mp];
}

// compile-time loop helper: f(integral_constant<int, LO>), ..., f(integral_constant<int, HI-1>)
template <int LO, int HI, typename F>
__device__ __forceinline__ void t32_for(F&& f)
{
    if constexpr (LO < HI) {
        f(std::integral_constant<int, LO>{});
        t32_for<LO + 1, HI>(f);
    }
}

// ---- check node: D inputs, D leave-one-out outputs, 4 frames (kernels_template_irreg.cl:60-96 / :205-245) ------------
template <int D>
__device__ __forceinline__ void cn_word_t32(const uint32_t (&w)[D], uint32_t (&o)[D], const uint8_t* tab, uint32_t lane4)
{
    constexpr int NST = D - 2;
    constexpr bool kPlain = t32_striped(NST) < NST;
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        uint32_t b[D], ms[D], mp[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            b[k] = __byte_perm(w[k], 0u, 0x4440u + f);
            ms[k] = ((b[k] & 0x1Cu) << 5) | (b[k] & 3u) | lane4;
            mp[k] = kPlain ? b[k] : 0u;
        }
        uint32_t P[D > 1 ? D : 2];
        P[1] = b[0];
        t32_for<1, D - 1>([&](auto J) {
            constexpr int j = decltype(J)::value;
            P[j + 1] = t32_lut<NST, j - 1>(tab, P[j], ms[j], mp[j]);
        });
        t32_for<0, D>([&](auto WO) {
            constexpr int wo = decltype(WO)::value;
            uint32_t t = (wo == 0) ? b[1] : P[wo];
            t32_for<(wo == 0 ? 2 : wo + 1), D>([&](auto K) {
                constexpr int k = decltype(K)::value;
                t = t32_lut<NST, k - 2>(tab, t, ms[k], mp[k]);
            });
            o[wo] = put_byte(o[wo], t, f);
        });
    }
}

// ---- variable node: channel value + D inbox messages, 4 frames (:125-177, decision :277-300) --------------------------
template <int D, bool DECIDE>
__device__ __forceinline__ void vn_word_t32(uint32_t chw, const uint32_t (&w)[D], uint32_t (&o)[D], uint32_t& dec,
                                            const uint8_t* tab, uint32_t lane4)
{
    constexpr int NST = DECIDE ? D : D - 1;
    constexpr bool kPlain = t32_striped(NST) < NST;
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0;
    dec = 0;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        uint32_t ms[D + 1], mp[D + 1];   // offsets of y_k, k = 1..D
#pragma unroll
        for (int k = 1; k <= D; ++k) {
            const uint32_t b = __byte_perm(w[k - 1], 0u, 0x4440u + f);
            ms[k] = ((b & 0x1Cu) << 5) | (b & 3u) | lane4;
            mp[k] = kPlain ? b : 0u;
        }
        uint32_t P[D + 2];
        P[1] = __byte_perm(chw, 0u, 0x4440u + f);
        t32_for<1, D>([&](auto J) {
            constexpr int j = decltype(J)::value;
            P[j + 1] = t32_lut<NST, j - 1>(tab, P[j], ms[j], mp[j]);
        });
        if constexpr (DECIDE) {
            dec = put_byte(dec, t32_lut<NST, D - 1>(tab, P[D], ms[D], mp[D]), f);
        } else {
            t32_for<1, D + 1>([&](auto WO) {
                constexpr int wo = decltype(WO)::value;
                uint32_t t = P[wo];
                t32_for<wo + 1, D + 1>([&](auto K) {
                    constexpr int k = decltype(K)::value;
                    t = t32_lut<NST, k - 2>(tab, t, ms[k], mp[k]);
                });
                o[wo - 1] = put_byte(o[wo - 1], t, f);
            });
        }
    }
}

// ---- one (node, tile) item ---------------------------------------------------------------------------------------
template <int MODE, bool EARLY, int D>
__device__ __forceinline__ uint32_t t32_item(const IbArgs& a, const uint8_t* tab, int node, int start, uint32_t col, uint32_t lane4)
{
    constexpr int VEC = t32_vec(MODE, D);
    if constexpr (MODE == kPhaseCn) {
        uint32_t m[D][VEC];
        if (a.iter0) {
#pragma unroll
            for (int k = 0; k < D; ++k) ld_words<VEC>(a.ch + (uint64_t)(uint32_t)a.vidx[start + k] * a.pitch + col, m[k]);
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) ld_words<VEC>(a.msg + (uint64_t)(uint32_t)(start + k) * a.pitch + col, m[k]);
        }
        uint32_t syn = 0;
        uint32_t r[D][VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            uint32_t w[D], o[D];
#pragma unroll
            for (int k = 0; k < D; ++k) w[k] = m[k][j];
            if (EARLY && !a.iter0) {
                // calc_syndrome (:304-325) on the VN->CN messages just read: parity of (msg < T/2), per frame byte
                uint32_t par = 0;
                if (a.tshift >= 0) {
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
                const int nv = a.B - (int)col - 4 * j;   // ignore padding frames
                const uint32_t vmask = nv >= 4 ? 0xffffffffu : nv <= 0 ? 0u : ((1u << (8 * nv)) - 1u);
                syn |= par & vmask;
            }
            cn_word_t32<D>(w, o, tab, lane4);
#pragma unroll
            for (int k = 0; k < D; ++k) r[k][j] = o[k];
        }
#pragma unroll
        for (int k = 0; k < D; ++k) st_words<VEC>(a.msg + (uint64_t)(uint32_t)(start + k) * a.pitch + col, r[k]);
        return syn;
    } else {
        constexpr bool DECIDE = MODE == kPhaseOut;
        constexpr bool kKeepRows = D <= 6;
        int rows[D];
        uint32_t c[VEC], m[D][VEC];
        ld_words<VEC>(a.ch + (uint64_t)(uint32_t)node * a.pitch + col, c);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            rows[k] = a.tv[start + k];
            ld_words<VEC>(a.msg + (uint64_t)(uint32_t)rows[k] * a.pitch + col, m[k]);
        }
        if (!DECIDE && D == 1) {   // degree-1 variable node forwards the raw channel value (:132-136)
            st_words<VEC>(a.msg + (uint64_t)(uint32_t)rows[0] * a.pitch + col, c);
            return 0;
        }
        uint32_t r[D][VEC], dec[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            uint32_t w[D], o[D];
#pragma unroll
            for (int k = 0; k < D; ++k) w[k] = m[k][j];
            vn_word_t32<D, DECIDE>(c[j], w, o, dec[j], tab, lane4);
            if (!DECIDE) {
#pragma unroll
                for (int k = 0; k < D; ++k) r[k][j] = o[k];
            }
        }
        if constexpr (DECIDE) {
            st_words<VEC>(a.out + (uint64_t)(uint32_t)node * a.pitch + col, dec);
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

template <int MODE, bool EARLY, int D>
__global__ void __launch_bounds__(kT32Threads, 1) ib_t32_kernel(T32Args p)
{
    constexpr int VEC = t32_vec(MODE, D);
    constexpr int NST = t32_cols(MODE, D);
    extern __shared__ __align__(128) uint8_t s_img[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ int s_next;
    const IbArgs& a = p.a;
    if (MODE != kPhaseOut && (EARLY || a.early) && a.it >= 1 && a.flags[a.it - 1] == 0) return;   // batch already converged
    const int lane = threadIdx.x & 31;
    const uint32_t lane4 = lane * 4;
    const uint32_t tl = (a.pitch + 128u * VEC - 1) / (128u * VEC);   // tiles of 32 lanes x 4 VEC frames per row
    const long long items = (long long)p.n_nodes * tl;
    const int lo = (int)(items * blockIdx.x / gridDim.x), hi = (int)(items * (blockIdx.x + 1) / gridDim.x);
    constexpr bool kHasTables = NST > 0;
    if (threadIdx.x == 0) {
        s_next = lo + kT32Threads / 32;
        if (kHasTables) {
            const uint8_t* img = p.image;
            if (MODE == kPhaseOut) img += (long long)executed_passes(a) * p.image_stride;
            phase_image_issue(s_img, img, (uint32_t)t32_image_bytes(NST), &s_mbar);
        }
        if (MODE == kPhaseOut && blockIdx.x == 0) *a.inum = executed_passes(a) + 1;
    }
    __syncthreads();
    uint32_t syn = 0;
    bool have_image = !kHasTables;
    int i = lo + (threadIdx.x >> 5);
    int node = 0, start = 0;
    if (i < hi) {
        const uint32_t ni = (uint32_t)i / tl;
        node = p.nodes[ni];
        start = p.starts[ni];
    }
    while (i < hi) {
        int i2 = 0;
        if (lane == 0) i2 = atomicAdd(&s_next, 1);
        i2 = __shfl_sync(0xffffffffu, i2, 0);
        int node2 = 0, start2 = 0;
        if (i2 < hi) {
            const uint32_t ni2 = (uint32_t)i2 / tl;
            node2 = p.nodes[ni2];
            start2 = p.starts[ni2];
        }
        const uint32_t tile = (uint32_t)i - ((uint32_t)i / tl) * tl;
        const uint32_t col = (tile * 32u + lane) * (4u * VEC);
        if (!have_image) {
            phase_image_wait(&s_mbar);
            have_image = true;
        }
        if (col < a.pitch) syn |= t32_item<MODE, EARLY, D>(a, s_img, node, start, col, lane4);
        i = i2;
        node = node2;
        start = start2;
    }
    if (MODE == kPhaseCn && EARLY && !a.iter0) {
        const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
        if (any != 0 && lane == 0) atomicOr(&a.flags[a.it], 1);
    }
    if (!have_image) phase_image_wait(&s_mbar);
}

// ---- fused per-phase kernel: ONE launch per phase, the degree classes one after the other inside every CTA ---------
// The per-class launches above pay launch latency + a 32-200 KB image per class and leave a partial wave behind every
// class (802.11n: 6 launches per iteration, the four small ones 11-26 us each at B = 32768; 300 launches per decode).
// Here a CTA owns a contiguous share of the items of EVERY class and walks through the classes: block barrier, one thread
// issues the TMA copy of the next class's image into the same shared-memory buffer, block barrier, dynamic item loop of
// that class (t32_item<MODE, EARLY, D>, selected by a switch over the class degree).  CTAs do not wait for each other, so
// nothing but a CTA's own last item sits between two classes.  Any degree set of up to kT32MaxClasses classes.
constexpr int kT32MaxClasses = 4;
struct T32PhaseArgs {
    IbArgs a;
    int n_cls;
    int deg[kT32MaxClasses];                    // heaviest class first
    const uint8_t* image[kT32MaxClasses];       // table image of every class for this launch
    long long image_stride[kT32MaxClasses];     // decision phase: bytes between the images of consecutive iterations
    const int* nodes[kT32MaxClasses];
    const int* starts[kT32MaxClasses];
    int n_nodes[kT32MaxClasses];
};

__device__ __forceinline__ void t32_wait_parity(uint64_t* mbar, uint32_t parity)
{
    const uint32_t mb = smem_u32(mbar);
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
}

// item loop of one class inside the fused kernel: items [., hi) of this CTA, first 32 taken statically, the rest from s_next
template <int MODE, bool EARLY, int D>
__device__ __forceinline__ uint32_t t32_class_items(const IbArgs& a, const uint8_t* s_img, const int* __restrict__ nodes,
                                                    const int* __restrict__ starts, int lo, int hi, int* s_next, uint64_t* mbar,
                                                    uint32_t parity, bool tables)
{
    constexpr int VEC = t32_vec(MODE, D);
    const int lane = threadIdx.x & 31;
    const uint32_t lane4 = lane * 4;
    const uint32_t tl = (a.pitch + 128u * VEC - 1) / (128u * VEC);
    uint32_t syn = 0;
    bool have_image = !tables;
    int i = lo + (threadIdx.x >> 5);
    int node = 0, start = 0;
    if (i < hi) {
        const uint32_t ni = (uint32_t)i / tl;
        node = nodes[ni];
        start = starts[ni];
    }
    while (i < hi) {
        int i2 = 0;
        if (lane == 0) i2 = atomicAdd(s_next, 1);
        i2 = __shfl_sync(0xffffffffu, i2, 0);
        int node2 = 0, start2 = 0;
        if (i2 < hi) {
            const uint32_t ni2 = (uint32_t)i2 / tl;
            node2 = nodes[ni2];
            start2 = starts[ni2];
        }
        const uint32_t tile = (uint32_t)i - ((uint32_t)i / tl) * tl;
        const uint32_t col = (tile * 32u + lane) * (4u * VEC);
        if (!have_image) {
            t32_wait_parity(mbar, parity);
            have_image = true;
        }
        if (col < a.pitch) syn |= t32_item<MODE, EARLY, D>(a, s_img, node, start, col, lane4);
        i = i2;
        node = node2;
        start = start2;
    }
    if (!have_image) t32_wait_parity(mbar, parity);   // every warp observes every copy: the barrier is re-armed for the next class
    return syn;
}

// the classes of one phase, one after the other (shared by the per-phase and the cooperative whole-decode kernel);
// the image of class c is p.image[c] + index * p.image_stride[c] (per-phase kernels: index = 0 except for the decision phase,
// where it is the number of executed passes).  Returns the syndrome bits seen (check-node phase with EARLY).
template <int MODE, bool EARLY>
__device__ __forceinline__ uint32_t t32_phase_classes(const IbArgs& a, const T32PhaseArgs& p, int index, uint8_t* s_img, int* s_next,
                                                      uint64_t* s_mbar, uint32_t& parity)
{
    const uint32_t mb = smem_u32(s_mbar);
    uint32_t syn = 0;
    for (int c = 0; c < p.n_cls; ++c) {
        const int d = p.deg[c];
        const int nst = t32_cols(MODE, d);
        const uint32_t vec = (uint32_t)t32_vec(MODE, d);
        const uint32_t tl = (a.pitch + 128u * vec - 1) / (128u * vec);
        const long long items = (long long)p.n_nodes[c] * tl;
        const int lo = (int)(items * blockIdx.x / gridDim.x), hi = (int)(items * (blockIdx.x + 1) / gridDim.x);
        if (hi <= lo) continue;              // uniform per CTA: nothing of this class here, no image needed
        const bool tables = nst > 0;
        __syncthreads();                     // every warp has left the previous image and counter (first class: mbarrier initialised)
        if (threadIdx.x == 0) {
            *s_next = lo + kT32Threads / 32;
            if (tables) {
                const uint8_t* img = p.image[c] + (long long)index * p.image_stride[c];
                const uint32_t bytes = (uint32_t)t32_image_bytes(nst);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the old image before the async writes
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
        }
        __syncthreads();
#define IBLDPC_T32_CASE(DD) \
    case DD: syn |= t32_class_items<MODE, EARLY, DD>(a, s_img, p.nodes[c], p.starts[c], lo, hi, s_next, s_mbar, parity, tables); break;
        if constexpr (MODE == kPhaseCn) {
            switch (d) {
                IBLDPC_T32_CASE(3) IBLDPC_T32_CASE(4) IBLDPC_T32_CASE(5) IBLDPC_T32_CASE(6)
                IBLDPC_T32_CASE(7) IBLDPC_T32_CASE(8) IBLDPC_T32_CASE(9) IBLDPC_T32_CASE(10)
            default: break;
            }
        } else {
            switch (d) {
                IBLDPC_T32_CASE(1) IBLDPC_T32_CASE(2) IBLDPC_T32_CASE(3) IBLDPC_T32_CASE(4) IBLDPC_T32_CASE(5) IBLDPC_T32_CASE(6)
                IBLDPC_T32_CASE(7) IBLDPC_T32_CASE(8) IBLDPC_T32_CASE(9) IBLDPC_T32_CASE(10) IBLDPC_T32_CASE(11) IBLDPC_T32_CASE(12)
            default: break;
            }
        }
#undef IBLDPC_T32_CASE
        if (tables) parity ^= 1u;
    }
    return syn;
}

template <int MODE, bool EARLY>
__global__ void __launch_bounds__(kT32Threads, 1) ib_t32_phase_kernel(T32PhaseArgs p)
{
    extern __shared__ __align__(128) uint8_t s_img[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ int s_next;
    const IbArgs& a = p.a;
    if (MODE != kPhaseOut && (EARLY || a.early) && a.it >= 1 && a.flags[a.it - 1] == 0) return;   // batch already converged
    int passes = 0;
    if (threadIdx.x == 0) {
        const uint32_t mb = smem_u32(&s_mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (MODE == kPhaseOut && blockIdx.x == 0) *a.inum = executed_passes(a) + 1;
    }
    if (MODE == kPhaseOut) passes = executed_passes(a);
    uint32_t parity = 0;
    const uint32_t syn = t32_phase_classes<MODE, EARLY>(a, p, passes, s_img, &s_next, &s_mbar, parity);
    if (MODE == kPhaseCn && EARLY && !a.iter0) {
        const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
        if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[a.it], 1);
    }
}

// ---- small batches: the whole decode in ONE cooperative launch (the |T| <= 32 counterpart of ib_coop_phase_kernel) ----
// Same class walk per phase, a grid-wide barrier between the phases instead of a kernel boundary: at the reference's
// 802.11n msg_at_time = 2000 a decode is 100 phases of ~10 us, most of it launch latency.
struct T32CoopArgs {
    IbArgs a;
    T32PhaseArgs cn;    // image[k] = image of table block 0 of class k, image_stride[k] = bytes between consecutive blocks
    T32PhaseArgs vn;    // update images of iteration 0 / stride
    T32PhaseArgs out;   // decision images of iteration 0 / stride (same classes as vn)
};

template <bool EARLY>
__global__ void __launch_bounds__(kT32Threads, 1) ib_t32_coop_kernel(T32CoopArgs p)
{
    namespace cg = cooperative_groups;
    extern __shared__ __align__(128) uint8_t s_img[];
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ int s_next;
    cg::grid_group grid = cg::this_grid();
    const IbArgs& a = p.a;
    if (threadIdx.x == 0) {
        const uint32_t mb = smem_u32(&s_mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t parity = 0;
    auto cn_phase = [&](int it) {
        IbArgs b = a;
        b.it = it; b.iter0 = (it < 0);
        const uint32_t syn = t32_phase_classes<kPhaseCn, EARLY>(b, p.cn, it + 1, s_img, &s_next, &s_mbar, parity);
        if (EARLY && it >= 0) {
            const unsigned any = __ballot_sync(0xffffffffu, syn != 0);
            if (any != 0 && (threadIdx.x & 31) == 0) atomicOr(&a.flags[it], 1);
        }
    };
    auto vn_phase = [&](int it) {
        IbArgs b = a;
        b.it = it; b.iter0 = 0;
        t32_phase_classes<kPhaseVn, false>(b, p.vn, it, s_img, &s_next, &s_mbar, parity);
    };
    cn_phase(-1);
    int passes = 0;
    for (int it = 0; it < a.imax - 1; ++it) {
        grid.sync();
        // reference stop rule (discrete_LDPC_decoder.py:233-276); flags[] were written before the grid barrier
        if (EARLY && it >= 1 && *reinterpret_cast<volatile int*>(&a.flags[it - 1]) == 0) break;
        vn_phase(it);
        grid.sync();
        cn_phase(it);
        passes = it + 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.inum = passes + 1;
    grid.sync();
    {
        IbArgs b = a;
        b.it = passes; b.iter0 = 0;
        t32_phase_classes<kPhaseOut, false>(b, p.out, passes, s_img, &s_next, &s_mbar, parity);
    }
}

using T32PhaseKernel = void (*)(T32PhaseArgs);
T32PhaseKernel t32_phase_cn_kernel(bool early);   // ib_t32_phase.cu
T32PhaseKernel t32_phase_vn_kernel();
T32PhaseKernel t32_phase_out_kernel();
using T32CoopKernel = void (*)(T32CoopArgs);
T32CoopKernel t32_coop_kernel(bool early);

using T32Kernel = void (*)(T32Args);
T32Kernel t32_cn_kernel(int d, bool early);   // ib_t32_cn.cu, d in [3, 10]
T32Kernel t32_vn_kernel(int d);               // ib_t32_vn.cu, d in [1, 12]
T32Kernel t32_out_kernel(int d);              // ib_t32_out.cu, d in [1, 12]
constexpr int kT32MaxDc = 10, kT32MaxDv = 12;

}  // namespace ibldpc
