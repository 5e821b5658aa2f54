// llr_kernels.cuh -- min-sum / belief-propagation benchmark decoders, channel quantizer,
// direct (inversion-method) sampling and error counting for sm_100a.
//
// Replaces Continous_LDPC_Decoding/kernels_min_and_BP.cl and
// AWGN_Channel_Transmission/kernels_quanti_template.cl (reference file:line cited per kernel).
//
// LLR message layout: the reference's two inboxes are kept (the stop rule needs the CN->VN
// messages of the last pass intact while the VN->CN ones are checked):
//   cin [n_edge][pitch]  check-node-major VN->CN messages, vin [n_edge][pitch] variable-node-major
//   CN->VN messages, element type F = float (fast) or double (bit-faithful to the reference).
// One warp = one (node, tile of 32 x VEC frames), VEC = 16 / sizeof(F): 128-bit accesses.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "ib_kernels.cuh"

namespace ibldpc {

struct LlrArgs {
    const int* __restrict__ sc;
    const int* __restrict__ deg_c;
    const int* __restrict__ tc;   // CN-major slot -> VN-major row
    const int* __restrict__ sv;
    const int* __restrict__ deg_v;
    const int* __restrict__ tv;   // VN-major slot -> CN-major row
    int n_var, n_chk;
    const void* ch;
    void* cin;
    void* vin;
    void* out;
    long long pitch;   // elements per row, multiple of VEC
    int B;
    int tiles;
    int* flags;
    int* inum;
    int it, early, imax;
    const int* __restrict__ vidx;   // variable index of every CN-major slot (layered schedule, llr_layered.cu)
};

template <typename F> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<double> { static constexpr int N = 2; };

template <typename F> struct alignas(16) Vec { F v[VecOf<F>::N]; };

template <typename F> __device__ __forceinline__ Vec<F> vload(const F* p) { return *reinterpret_cast<const Vec<F>*>(p); }
template <typename F> __device__ __forceinline__ void vstore(F* p, const Vec<F>& x) { *reinterpret_cast<Vec<F>*>(p) = x; }

// sign(x)*min(150, sign(x)*x) with OpenCL sign(+-0) = 0  (kernels_min_and_BP.cl:8,69,120)
template <typename F> __device__ __forceinline__ F clip150(F x)
{
    return x > F(150) ? F(150) : (x < F(-150) ? F(-150) : (x == F(0) ? F(0) : x));
}

// boxplus (kernels_min_and_BP.cl:5-9).  double: the reference expression verbatim in meaning;
// float: the algebraically identical overflow-free form
//   sign(a)sign(b)min(|a|,|b|) + log1p(e^-|a+b|) - log1p(e^-|a-b|).
__device__ __forceinline__ double boxplus(double a, double b)
{
    const double bp = log((1.0 + exp(a + b)) / (exp(a) + exp(b)));
    return clip150(bp);
}
// Box-plus of the forward/backward recursion (ALGO 2).  float64: the reference expression itself -- inputs are clipped
// to +-150, so exp(a + b) <= e^300 cannot overflow; measured on B200 it is 1.46x faster than the log1p form
// (0.169 vs 0.115 Gbit/s, (3,6) n=8000): two double-precision log1p cost more than one log, one exp and one division.
__device__ __forceinline__ double boxplus_stable(double a, double b) { return boxplus(a, b); }
__device__ __forceinline__ float boxplus_stable(float a, float b);
__device__ __forceinline__ float boxplus(float a, float b)
{
    const float s = fminf(fabsf(a), fabsf(b));
    const float sg = ((a < 0.f) != (b < 0.f)) ? -s : s;
    // MUFU-based: e^-x in (0,1], so 1 + e^-x in (1,2] where __logf is accurate to ~1e-7 absolute
    const float bp = sg + __logf(1.0f + __expf(-fabsf(a + b))) - __logf(1.0f + __expf(-fabsf(a - b)));
    return clip150((a == 0.f || b == 0.f) ? 0.f : bp);
}

__device__ __forceinline__ float boxplus_stable(float a, float b) { return boxplus(a, b); }

// ---- check node ---------------------------------------------------------------------------
// ALGO 0: checknode_update_minsum (kernels_min_and_BP.cl:126-167).  The sequential
// t = sign(m*t)*min(|t|,|m|) equals prod(sign)*min(|.|) exactly (min is exact, any zero input
// gives +0), so min1/min2/argmin + sign parity reproduce it bit for bit.
// ALGO 1: checknode_update (BP, :32-71): sequential boxplus in slot order, shared prefixes.
template <typename F, int ALGO, int D>
__device__ __forceinline__ void llr_cn_compute(const Vec<F>* m, Vec<F>* o, int d)
{
    constexpr int V = VecOf<F>::N;
    const int dd = D > 0 ? D : d;
    if (ALGO == 0) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            F min1 = sizeof(F) == 4 ? F(3.0e38f) : F(1e300), min2 = min1;
            int arg = -1, neg = 0;
#pragma unroll
            for (int k = 0; k < dd; ++k) {
                const F x = m[k].v[e];
                const F ax = x < F(0) ? -x : x;
                neg ^= (x < F(0));
                if (ax < min1) { min2 = min1; min1 = ax; arg = k; }
                else if (ax < min2) { min2 = ax; }
            }
#pragma unroll
            for (int k = 0; k < dd; ++k) {
                const F x = m[k].v[e];
                F r;
                if (dd == 2) {
                    r = m[1 - k].v[e];
                } else {
                    const F mag = (k == arg) ? min2 : min1;
                    const int sneg = neg ^ (x < F(0));
                    r = mag == F(0) ? F(0) : (sneg ? -mag : mag);
                }
                o[k].v[e] = r;
            }
        }
    } else if (ALGO == 3 && sizeof(F) == 8 && dd >= 4) {
        // float64 default: the forward/backward recursion in the LIKELIHOOD-RATIO domain.  With A = e^a, B = e^b the
        // reference box-plus is log(R), R = (1 + A B) / (A + B) (kernels_min_and_BP.cl:7), and the next box-plus needs e^{log R}
        // = R again: carrying R instead of log R through the chains leaves d exp (one per input), d log (one per output)
        // and 3(d-2) divisions per check instead of 9(d-2) exp, 3(d-2) log and 3(d-2) divisions -- the kernel is bound by
        // the double-precision pipe.  The per-operation clip to +-150 (:8) is the clamp of R to [e^-150, e^150]; inputs
        // beyond |709| overflow exactly where the reference's exp(a + b) does.  Differences to the reference order are
        // rounding only (1e-13 on the golden vectors); the hard decisions stay identical (tests/test_gpu_parity.py:
        // >= 99.99 % of 20000 frames).  IBLDPC_BP_LOGDOMAIN=1 selects ALGO 2, IBLDPC_BP_SEQUENTIAL=1 the reference order.
        const double kRmax = 1.3937095806663797e65, kRmin = 7.1750959731644115e-66;   // e^150, e^-150
        auto rbox = [&](double A, double B) { return fmin(fmax((1.0 + A * B) / (A + B), kRmin), kRmax); };
#pragma unroll
        for (int e = 0; e < V; ++e) {
            double M[D > 0 ? D : kMaxGenericDeg], fw[D > 0 ? D : kMaxGenericDeg], bw[D > 0 ? D : kMaxGenericDeg];
#pragma unroll
            for (int k = 0; k < dd; ++k) M[k] = exp((double)m[k].v[e]);
            fw[0] = M[0];
#pragma unroll
            for (int k = 1; k <= dd - 2; ++k) fw[k] = rbox(M[k], fw[k - 1]);
            bw[dd - 1] = M[dd - 1];
#pragma unroll
            for (int k = dd - 2; k >= 1; --k) bw[k] = rbox(M[k], bw[k + 1]);
            o[0].v[e] = (F)clip150(log(bw[1]));
            o[dd - 1].v[e] = (F)clip150(log(fw[dd - 2]));
#pragma unroll
            for (int k = 1; k <= dd - 2; ++k) o[k].v[e] = (F)clip150(log(rbox(fw[k - 1], bw[k + 1])));
        }
    } else if ((sizeof(F) == 4 || ALGO == 2 || ALGO == 3) && dd >= 4) {
        // forward/backward box-plus, 3(d-2) operations instead of 2(d-2) + (d-1)(d-2)/2.  Box-plus is associative in
        // exact arithmetic; the result differs from the reference's sequential order only by rounding: fp32 (covered
        // by the fp32 tolerance) and -- ALGO 2, the float64 default -- float64, where the hard decisions stay identical
        // (tests/test_gpu_parity.py: >= 99.99 % of 20000 frames; IBLDPC_BP_SEQUENTIAL=1 restores the reference order).
#pragma unroll
        for (int e = 0; e < V; ++e) {
            F fw[D > 0 ? D : kMaxGenericDeg], bw[D > 0 ? D : kMaxGenericDeg];
            fw[0] = m[0].v[e];
#pragma unroll
            for (int k = 1; k <= dd - 2; ++k) fw[k] = (ALGO >= 2 ? boxplus_stable(m[k].v[e], fw[k - 1]) : boxplus(m[k].v[e], fw[k - 1]));
            bw[dd - 1] = m[dd - 1].v[e];
#pragma unroll
            for (int k = dd - 2; k >= 1; --k) bw[k] = (ALGO >= 2 ? boxplus_stable(m[k].v[e], bw[k + 1]) : boxplus(m[k].v[e], bw[k + 1]));
            o[0].v[e] = clip150(bw[1]);
            o[dd - 1].v[e] = clip150(fw[dd - 2]);
#pragma unroll
            for (int k = 1; k <= dd - 2; ++k) o[k].v[e] = clip150(ALGO >= 2 ? boxplus_stable(fw[k - 1], bw[k + 1]) : boxplus(fw[k - 1], bw[k + 1]));
        }
    } else {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            F P[D > 0 ? D + 1 : kMaxGenericDeg + 1];
            P[1] = m[0].v[e];
#pragma unroll
            for (int j = 1; j <= dd - 2; ++j) P[j + 1] = boxplus(m[j].v[e], P[j]);
#pragma unroll
            for (int wo = 0; wo < dd; ++wo) {
                F t = (wo == 0) ? m[1].v[e] : P[wo];
#pragma unroll
                for (int k = (wo == 0 ? 2 : wo + 1); k < dd; ++k) t = boxplus(m[k].v[e], t);
                o[wo].v[e] = clip150(t);
            }
        }
    }
}

// Low-batch min-sum check-node kernel (the reference runs DVB-S2 / WLAN min-sum with msg_at_time = 2,
// Irregular_LDPC_Decoding/DVB-S2/BER_simulation_OpenCL_min_sum.py): with only a handful of frames the
// frame-per-lane mapping leaves most lanes idle, so here the lanes of a warp are the EDGES of
// 32/seg check nodes (seg = power of two >= d_c_max): min / argmin / second min by warp-shuffle
// butterflies inside each seg-lane segment, the sign parity by a warp ballot.  Same results as
// checknode_update_minsum (kernels_min_and_BP.cl:126-167): the first minimum (lowest slot) wins ties.
template <typename F>
__global__ void __launch_bounds__(kThreads) llr_cn_minsum_shfl_kernel(LlrArgs a, int seg_log2)
{
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;
    const int seg = 1 << seg_log2;
    const int lane = threadIdx.x & 31;
    const int sub = lane >> seg_log2, e = lane & (seg - 1);
    const int per_warp = 32 >> seg_log2;
    const long long gwarp = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const F* cin = static_cast<const F*>(a.cin);
    F* vin = static_cast<F*>(a.vin);
    const F inf = sizeof(F) == 4 ? F(3.0e38f) : F(1e300);
    const unsigned segmask = (seg == 32 ? 0xffffffffu : ((1u << seg) - 1u)) << (sub * seg);
    for (long long base = gwarp * per_warp; base < a.n_chk; base += nwarps * per_warp) {
        const int c = (int)base + sub;
        const int d = c < a.n_chk ? a.deg_c[c] : 0;
        const bool valid = e < d;
        const int slot = valid ? a.sc[c] + e : 0;
        const int dst = valid ? a.tc[slot] : 0;
        for (int f = 0; f < a.B; ++f) {
            const F x = valid ? cin[(long long)slot * a.pitch + f] : inf;
            const F ax = x < F(0) ? -x : x;
            const unsigned negb = __ballot_sync(0xffffffffu, valid && x < F(0)) & segmask;
            F m1 = ax;
            int idx = e;
            for (int off = seg >> 1; off > 0; off >>= 1) {
                const F om = __shfl_xor_sync(0xffffffffu, m1, off);
                const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
                if (om < m1 || (om == m1 && oi < idx)) { m1 = om; idx = oi; }
            }
            F m2 = (e == idx) ? inf : ax;
            for (int off = seg >> 1; off > 0; off >>= 1) {
                const F om = __shfl_xor_sync(0xffffffffu, m2, off);
                m2 = om < m2 ? om : m2;
            }
            const F other = __shfl_xor_sync(0xffffffffu, x, 1);   // degree-2 checks forward the other input as is
            if (valid) {
                F r;
                if (d == 2) {
                    r = other;
                } else {
                    const F mag = (e == idx) ? m2 : m1;
                    const int sneg = (__popc(negb) & 1) ^ (x < F(0) ? 1 : 0);
                    r = mag == F(0) ? F(0) : (sneg ? -mag : mag);
                }
                vin[(long long)dst * a.pitch + f] = r;
            }
        }
    }
}

template <typename F, int ALGO, int D>
__global__ void __launch_bounds__(kThreads) llr_cn_kernel(LlrArgs a, const int* __restrict__ nodes, int n_nodes)
{
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;
    constexpr int V = VecOf<F>::N;
    constexpr int MAXD = D > 0 ? D : kMaxGenericDeg;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const long long items = (long long)n_nodes * a.tiles;
    const F* cin = static_cast<const F*>(a.cin);
    F* vin = static_cast<F*>(a.vin);
    for (long long item = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); item < items; item += nwarps) {
        const int i = (int)(item / a.tiles);
        const int tile = (int)(item - (long long)i * a.tiles);
        const long long col = ((long long)tile * 32 + lane) * V;
        if (col >= a.pitch) continue;
        const int c = nodes[i];
        const int s = a.sc[c];
        const int d = D > 0 ? D : a.deg_c[c];
        Vec<F> m[MAXD], o[MAXD];
#pragma unroll
        for (int k = 0; k < (D > 0 ? D : d); ++k) m[k] = vload(cin + (long long)(s + k) * a.pitch + col);
        llr_cn_compute<F, ALGO, D>(m, o, d);
#pragma unroll
        for (int k = 0; k < (D > 0 ? D : d); ++k) vstore(vin + (long long)a.tc[s + k] * a.pitch + col, o[k]);
    }
}

// ---- variable node ------------------------------------------------------------------------
// varnode_update (kernels_min_and_BP.cl:76-123): sequential sum in slot order, skipping the
// target edge, clipped to +-150.  MODE 0: update; MODE 1: calc_varnode_output (:170-204),
// unclipped sum of everything; MODE 2: send_channel_values_to_checknode_inbox (:12-29).
template <typename F, int MODE, int D>
__global__ void __launch_bounds__(kThreads) llr_vn_kernel(LlrArgs a, const int* __restrict__ nodes, int n_nodes)
{
    constexpr int V = VecOf<F>::N;
    constexpr int MAXD = D > 0 ? D : kMaxGenericDeg;
    __shared__ int s_passes;
    if (MODE == 1) {
        if (threadIdx.x == 0) {
            int passes = a.imax - 1;
            if (a.early)
                for (int it = 0; it < a.imax - 1; ++it)
                    if (a.flags[it] == 0) { passes = it + 1; break; }
            s_passes = passes;
            if (blockIdx.x == 0) *a.inum = passes + 1;
        }
        __syncthreads();
    } else if (MODE == 0 && a.early && a.it >= 1 && a.flags[a.it - 1] == 0) {
        return;
    }
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const long long items = (long long)n_nodes * a.tiles;
    const F* ch = static_cast<const F*>(a.ch);
    const F* vin = static_cast<const F*>(a.vin);
    F* cin = static_cast<F*>(a.cin);
    F* out = static_cast<F*>(a.out);
    for (long long item = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); item < items; item += nwarps) {
        const int i = (int)(item / a.tiles);
        const int tile = (int)(item - (long long)i * a.tiles);
        const long long col = ((long long)tile * 32 + lane) * V;
        if (col >= a.pitch) continue;
        const int v = nodes[i];
        const int s = a.sv[v];
        const int d = D > 0 ? D : a.deg_v[v];
        const Vec<F> c = vload(ch + (long long)v * a.pitch + col);
        if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < (D > 0 ? D : d); ++k) vstore(cin + (long long)a.tv[s + k] * a.pitch + col, c);
            continue;
        }
        Vec<F> m[MAXD];
#pragma unroll
        for (int k = 0; k < (D > 0 ? D : d); ++k) m[k] = vload(vin + (long long)(s + k) * a.pitch + col);
        if (MODE == 1) {
            Vec<F> r;
#pragma unroll
            for (int e = 0; e < V; ++e) {
                F t = c.v[e] + m[0].v[e];
#pragma unroll
                for (int k = 1; k < (D > 0 ? D : d); ++k) t = t + m[k].v[e];
                r.v[e] = t;
            }
            vstore(out + (long long)v * a.pitch + col, r);
        } else {
            Vec<F> o[MAXD];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                F P = c.v[e];   // running prefix ch + m_0 + ... + m_{w-1}
#pragma unroll
                for (int wo = 0; wo < (D > 0 ? D : d); ++wo) {
                    F t = P;
#pragma unroll
                    for (int k = wo + 1; k < (D > 0 ? D : d); ++k) t = t + m[k].v[e];
                    o[wo].v[e] = clip150(t);
                    P = P + m[wo].v[e];
                }
            }
#pragma unroll
            for (int k = 0; k < (D > 0 ? D : d); ++k) vstore(cin + (long long)a.tv[s + k] * a.pitch + col, o[k]);
        }
    }
}

// calc_syndrome (kernels_min_and_BP.cl:206-227) + the batch-wide sum
// (min_sum_decoder_irreg.py:269): any unsatisfied check of a valid frame raises flags[it].
template <typename F>
__global__ void __launch_bounds__(kThreads) llr_syndrome_kernel(LlrArgs a)
{
    if (a.it >= 1 && a.flags[a.it - 1] == 0) return;
    constexpr int V = VecOf<F>::N;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const long long items = (long long)a.n_chk * a.tiles;
    const F* cin = static_cast<const F*>(a.cin);
    bool bad = false;
    for (long long item = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); item < items; item += nwarps) {
        const int c = (int)(item / a.tiles);
        const int tile = (int)(item - (long long)c * a.tiles);
        const long long col = ((long long)tile * 32 + lane) * V;
        if (col >= a.pitch) continue;
        const int s = a.sc[c], d = a.deg_c[c];
        int par[V];
#pragma unroll
        for (int e = 0; e < V; ++e) par[e] = 0;
        for (int k = 0; k < d; ++k) {
            const Vec<F> x = vload(cin + (long long)(s + k) * a.pitch + col);
#pragma unroll
            for (int e = 0; e < V; ++e) par[e] ^= (x.v[e] < F(0));
        }
#pragma unroll
        for (int e = 0; e < V; ++e) bad |= (par[e] != 0) && (col + e < a.B);
    }
    const unsigned any = __ballot_sync(0xffffffffu, bad);
    if (any != 0 && lane == 0) atomicOr(&a.flags[a.it], 1);
}

// ---- quantizer ----------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), counter = element index, key = seed.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t index)
{
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t x = (((uint64_t)c[0] << 32) | c[1]) >> 11;
    return (double)x * (1.0 / 9007199254740992.0);   // 2^-53, u in [0,1)
}

// quantize / quantize_LLR (kernels_quanti_template.cl:2-27, :29-52):
//   cluster = #{w in [1,card) : x - limits[w] > 0}.  SRC 0: x from memory; SRC 1: x = Philox uniform
// (quantize_direct_OpenCL, AWGN_Quantizer_BPSK.py:201-228, without the host RNG + H2D copy).
// OUT 0: uint8 cluster, 1: float LLR, 2: double LLR, 3: the uniform itself (double).
template <int SRC, int OUT>
__global__ void quantize_kernel(const double* __restrict__ x, long long n, const double* __restrict__ limits, int card,
                                const double* __restrict__ llr, uint64_t seed, uint64_t offset, void* out)
{
    extern __shared__ double s_q[];   // limits[card] then llr[card]
    for (int i = threadIdx.x; i < card; i += blockDim.x) {
        s_q[i] = limits ? limits[i] : 0.0;
        if (OUT == 1 || OUT == 2) s_q[card + i] = llr[i];
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double xv = SRC == 0 ? x[i] : philox_uniform(seed, offset + (uint64_t)i);
        if (OUT == 3) { static_cast<double*>(out)[i] = xv; continue; }
        int cl = 0;
        for (int w = 1; w != card; ++w) cl += ((xv - s_q[w]) > 0.0) ? 1 : 0;
        if (OUT == 0) static_cast<uint8_t*>(out)[i] = (uint8_t)cl;
        if (OUT == 1) static_cast<float*>(out)[i] = (float)s_q[card + cl];
        if (OUT == 2) static_cast<double*>(out)[i] = s_q[card + cl];
    }
}

// ---- error counting -------------------------------------------------------------------------
// return_errors_all_zero (discrete_LDPC_decoder_irreg.py:343-349 / min_sum_decoder_irreg.py:290-295):
// decided bit = (value < threshold); compared with 0 (all-zero codeword) or ref_bits.
// One thread per frame column and 64-row chunk; counters[0] += bit errors, frame_err[f] |= 1.
template <typename E>
__global__ void count_errors_kernel(const E* __restrict__ out, long long rows, long long B, E threshold,
                                    const uint8_t* __restrict__ ref_bits, unsigned long long* counters,
                                    int* frame_err)
{
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * 64;
    const long long r1 = r0 + 64 < rows ? r0 + 64 : rows;
    unsigned int e = 0;
    if (f < B) {
        for (long long r = r0; r < r1; ++r) {
            const unsigned bit = out[r * B + f] < threshold ? 1u : 0u;
            const unsigned ref = ref_bits ? (ref_bits[r * B + f] != 0) : 0u;
            e += bit ^ ref;
        }
        if (e) atomicOr(&frame_err[f], 1);
    }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e) atomicAdd(&counters[0], (unsigned long long)e);
}
static __global__ void count_frames_kernel(const int* __restrict__ frame_err, long long B, unsigned long long* counters)
{
    unsigned int e = 0;
    for (long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (long long)gridDim.x * blockDim.x)
        e += frame_err[f] != 0;
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((threadIdx.x & 31) == 0 && e) atomicAdd(&counters[1], (unsigned long long)e);
}

}  // namespace ibldpc
