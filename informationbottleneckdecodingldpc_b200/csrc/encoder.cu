// encoder.cu -- transmit side of the BER drivers on the GPU: systematic LDPC encoder, random
// information bits, BPSK + AWGN channel.  Part of libibldpc.so (C ABI in include/ibldpc.h).
//
// Replaces Discrete_LDPC_decoding/LDPC_encoder.py (+ GF2MatrixMul_c.pyx), the per-frame Python loop of
// AWGN_Channel_Transmission/LDPC_Transmitter.py:109-125 and AWGN_channel.py:32-50.
//
// The reference encodes like MATLAB's comm.LDPCEncoder: the codeword is [x ; p] with H_last p = H_first x over
// GF(2), H_last = the last N-K columns of H, solved by substitution when H_last (or its row reversal) is
// triangular and through a GF(2) factorisation otherwise (LDPC_encoder.py:86-123, :196-262).  Since H_last is
// invertible the parity vector is unique, so any exact solver gives the reference's codewords bit for bit.
//
// Data layout: frames are independent, so everything is done on 32 frames at once: bit f of word w of row r
// is the bit of row r in frame 32 w + f ("frame-packed" words, (rows, words) row-major, words padded to 4).
//   1. pack_bits_kernel      (K, B) uint8 information bits -> frame-packed words          (warp ballot)
//   2. gf2_spmv_kernel       s = H_first x                  one thread per (check row, word), CSR gather
//   3. gf2_trisolve_kernel   substitution in the order the host analysis found: step t solves variable var[t]
//                            from equation eq[t]; sequential in t, parallel over words (one thread per word)
//      gf2_dense_kernel      "Matrix Inverse" codes: p = G s with the dense M x M matrix G = H_last^-1 (bit-packed
//                            rows, computed once on the host by GF(2) elimination)
//   4. unpack_codeword_kernel  (N, B) uint8 codeword [x ; p]
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ibldpc.h"
#include "llr_kernels.cuh"   // philox4x32_10

namespace ibldpc {
int fail_msg(int code, const std::string& msg);   // ibldpc.cu: sets the thread's last-error string
}
using ibldpc::fail_msg;

#define ECK(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail_msg(IBLDPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)

namespace {

__global__ void pack_bits_kernel(const uint8_t* __restrict__ bits, uint32_t* __restrict__ words, int rows, long long B,
                                 int wp)
{
    // one warp per destination word: lane = frame inside the word, the word is the ballot of the bits
    const long long n_words = (long long)rows * wp;
    const int lane = threadIdx.x & 31;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_words;
         w += ((long long)gridDim.x * blockDim.x) >> 5) {
        const long long r = w / wp, c = w - r * wp;
        const long long f = c * 32 + lane;
        const unsigned bit = (f < B) ? (bits[r * B + f] & 1u) : 0u;
        const unsigned word = __ballot_sync(0xffffffffu, bit != 0);
        if (lane == 0) words[w] = word;
    }
}

// s[r] = XOR of x[c] over the entries c of CSR row r
__global__ void gf2_spmv_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const uint32_t* __restrict__ x,
                                uint32_t* __restrict__ s, int n_rows, int wp)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= wp) return;
    for (int r = blockIdx.y; r < n_rows; r += gridDim.y) {
        uint32_t acc = 0;
        for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) acc ^= x[(long long)col[e] * wp + w];
        s[(long long)r * wp + w] = acc;
    }
}

// Substitution: for t = 0..M-1:  p[var[t]] = s[eq[t]] ^ XOR_{e in [optr[t], optr[t+1])} p[oth[e]].
// Every p row read in step t was written by the same thread in an earlier step.  The recursion is a chain of
// short steps, so everything that does not depend on it -- the right-hand sides s[eq[t]] and the schedule entries
// of the next kDepth steps -- is fetched a block ahead (independent loads, one round trip per block instead of one
// per step); the parity bit of the previous step stays in a register (staircase codes read nothing else).
constexpr int kTriDepth = 16;
__global__ void gf2_trisolve_kernel(const int* __restrict__ eq, const int* __restrict__ var, const int* __restrict__ optr,
                                    const int* __restrict__ oth, const uint32_t* __restrict__ s, uint32_t* __restrict__ p,
                                    int M, int wp)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= wp) return;
    int prev_var = -1;
    uint32_t prev_val = 0;
    for (int t0 = 0; t0 < M; t0 += kTriDepth) {
        uint32_t rhs[kTriDepth];
        int v[kTriDepth], lo[kTriDepth + 1];
#pragma unroll
        for (int q = 0; q < kTriDepth; ++q) {
            const int t = min(t0 + q, M - 1);
            rhs[q] = s[(long long)eq[t] * wp + w];
            v[q] = var[t];
            lo[q] = optr[min(t0 + q, M)];      // optr has M + 1 entries
        }
        lo[kTriDepth] = optr[min(t0 + kTriDepth, M)];
        int j0[kTriDepth];   // first operand of every step of the block (-1 = none), fetched ahead as well
#pragma unroll
        for (int q = 0; q < kTriDepth; ++q) j0[q] = lo[q] < lo[q + 1] ? oth[lo[q]] : -1;
#pragma unroll
        for (int q = 0; q < kTriDepth; ++q) {
            if (t0 + q < M) {
                uint32_t acc = rhs[q];
                if (j0[q] >= 0) acc ^= (j0[q] == prev_var) ? prev_val : p[(long long)j0[q] * wp + w];
                for (int e = lo[q] + 1; e < lo[q + 1]; ++e) {
                    const int j = oth[e];
                    acc ^= (j == prev_var) ? prev_val : p[(long long)j * wp + w];
                }
                p[(long long)v[q] * wp + w] = acc;
                prev_var = v[q];
                prev_val = acc;
            }
        }
    }
}

// p[r] = XOR over the set bits k of G row r of s[k]  (G: M rows of gw 32-bit words)
__global__ void gf2_dense_kernel(const uint32_t* __restrict__ G, int gw, const uint32_t* __restrict__ s, uint32_t* __restrict__ p,
                                 int M, int wp4)
{
    const int w4 = blockIdx.x * blockDim.x + threadIdx.x;   // uint4 column
    if (w4 >= wp4) return;
    const uint4* s4 = reinterpret_cast<const uint4*>(s);
    uint4* p4 = reinterpret_cast<uint4*>(p);
    for (int r = blockIdx.y; r < M; r += gridDim.y) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int kw = 0; kw < gw; ++kw) {
            uint32_t g = G[(long long)r * gw + kw];   // warp-uniform
            while (g) {
                const int k = kw * 32 + __ffs(g) - 1;
                g &= g - 1;
                const uint4 v = s4[(long long)k * wp4 + w4];
                acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
            }
        }
        p4[(long long)r * wp4 + w4] = acc;
    }
}

__global__ void unpack_codeword_kernel(const uint8_t* __restrict__ bits, const uint32_t* __restrict__ p, uint8_t* __restrict__ cw,
                                       int K, int N, long long B, int wp)
{
    const long long n = (long long)N * B;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / B, f = i - r * B;
        cw[i] = r < K ? (bits[i] & 1u) : (uint8_t)((p[(r - K) * wp + (f >> 5)] >> (f & 31)) & 1u);
    }
}

// MODE 0: uint8 random bits (bit = top bit of the Philox word, counter = element index)
// MODE 1: y = x + sigma * n, x double;  MODE 2: y = (1 - 2 bit) + sigma * n, bit uint8 (BPSK mapping 0 -> +1, 1 -> -1,
// LDPC_Transmitter.py:127-132).  n ~ N(0,1) by Box-Muller from one Philox4x32-10 block per element.
template <int MODE>
__global__ void channel_kernel(const void* __restrict__ in, long long n, double sigma, uint64_t seed, uint64_t offset,
                               void* __restrict__ out)
{
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)(offset + (uint64_t)i), (uint32_t)((offset + (uint64_t)i) >> 32), 0u, 0u};
        ibldpc::philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        if (MODE == 0) {
            static_cast<uint8_t*>(out)[i] = (uint8_t)(c[0] >> 31);
        } else {
            const uint64_t a = (((uint64_t)c[0] << 32) | c[1]) >> 11, b = (((uint64_t)c[2] << 32) | c[3]) >> 11;
            const double u1 = ((double)a + 1.0) * (1.0 / 9007199254740992.0);   // (0, 1]
            const double u2 = (double)b * (1.0 / 9007199254740992.0);           // [0, 1)
            const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
            const double x = MODE == 1 ? static_cast<const double*>(in)[i]
                                       : 1.0 - 2.0 * (double)(static_cast<const uint8_t*>(in)[i] & 1u);
            static_cast<double*>(out)[i] = x + sigma * z;
        }
    }
}

template <typename T>
int upload_vec(T** dst, const T* src, size_t n)
{
    ECK(cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) ECK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return IBLDPC_OK;
}

}  // namespace

struct ibldpc_encoder {
    int device = 0;
    int N = 0, K = 0, M = 0, method = 0;
    int *d_a_rowptr = nullptr, *d_a_col = nullptr;                                   // H_first, CSR
    int *d_eq = nullptr, *d_var = nullptr, *d_optr = nullptr, *d_oth = nullptr;      // substitution schedule
    uint32_t* d_G = nullptr;                                                         // dense H_last^-1, bit-packed rows
    int gw = 0;
    uint32_t *d_x = nullptr, *d_s = nullptr, *d_p = nullptr;                         // frame-packed scratch
    size_t scratch_words = 0;
};

extern "C" {

int ibldpc_encoder_create(const ibldpc_encoder_desc* d, int device, ibldpc_encoder_handle* out)
{
    if (!d || !out) return fail_msg(IBLDPC_E_INVALID, "null argument");
    const int N = d->n_var, K = d->n_info, M = N - K;
    if (N <= 0 || K <= 0 || M <= 0) return fail_msg(IBLDPC_E_INVALID, "need 0 < n_info < n_var");
    if (!d->a_rowptr || !d->a_col) return fail_msg(IBLDPC_E_INVALID, "null H_first tables");
    if (d->a_rowptr[0] != 0) return fail_msg(IBLDPC_E_INVALID, "a_rowptr[0] must be 0");
    for (int r = 0; r < M; ++r)
        if (d->a_rowptr[r + 1] < d->a_rowptr[r]) return fail_msg(IBLDPC_E_INVALID, "a_rowptr is not monotone");
    for (int e = 0; e < d->a_rowptr[M]; ++e)
        if (d->a_col[e] < 0 || d->a_col[e] >= K) return fail_msg(IBLDPC_E_INVALID, "a_col entry outside [0, n_info)");
    if (d->method == IBLDPC_ENC_SUBSTITUTION) {
        if (!d->eq || !d->var || !d->oth_ptr || (!d->oth && d->oth_ptr[M] > 0))
            return fail_msg(IBLDPC_E_INVALID, "null substitution tables");
        // every variable solved exactly once, every equation used exactly once, operands solved earlier
        std::vector<int> solved_at(M, -1), used(M, 0);
        for (int t = 0; t < M; ++t) {
            const int v = d->var[t], e = d->eq[t];
            if (v < 0 || v >= M || e < 0 || e >= M || solved_at[v] >= 0 || used[e])
                return fail_msg(IBLDPC_E_INVALID, "eq / var are not permutations of the parity rows");
            for (int q = d->oth_ptr[t]; q < d->oth_ptr[t + 1]; ++q) {
                const int j = d->oth[q];
                if (j < 0 || j >= M || solved_at[j] < 0)
                    return fail_msg(IBLDPC_E_INVALID, "substitution step reads a parity bit that is not solved yet");
            }
            solved_at[v] = t;
            used[e] = 1;
        }
    } else if (d->method == IBLDPC_ENC_DENSE) {
        if (!d->dense_inverse) return fail_msg(IBLDPC_E_INVALID, "null dense inverse");
    } else {
        return fail_msg(IBLDPC_E_INVALID, "unknown encoding method");
    }
    int ndev = 0;
    ECK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail_msg(IBLDPC_E_INVALID, "no such CUDA device");
    ECK(cudaSetDevice(device));
    ibldpc_encoder* h = new ibldpc_encoder();
    h->device = device; h->N = N; h->K = K; h->M = M; h->method = d->method;
    int rc = upload_vec(&h->d_a_rowptr, d->a_rowptr, (size_t)M + 1);
    if (!rc) rc = upload_vec(&h->d_a_col, d->a_col, (size_t)d->a_rowptr[M]);
    if (!rc && d->method == IBLDPC_ENC_SUBSTITUTION) {
        rc = upload_vec(&h->d_eq, d->eq, (size_t)M);
        if (!rc) rc = upload_vec(&h->d_var, d->var, (size_t)M);
        if (!rc) rc = upload_vec(&h->d_optr, d->oth_ptr, (size_t)M + 1);
        if (!rc) rc = upload_vec(&h->d_oth, d->oth, (size_t)d->oth_ptr[M]);
    }
    if (!rc && d->method == IBLDPC_ENC_DENSE) {
        h->gw = (M + 31) / 32;
        rc = upload_vec(&h->d_G, d->dense_inverse, (size_t)M * h->gw);
    }
    if (rc) { ibldpc_encoder_destroy(h); return rc; }
    *out = h;
    return IBLDPC_OK;
}

int ibldpc_encode(ibldpc_encoder_handle h, const uint8_t* bits_dev, int64_t B, uint8_t* codeword_dev, void* stream)
{
    if (!h || !bits_dev || !codeword_dev) return fail_msg(IBLDPC_E_INVALID, "null argument");
    if (B <= 0 || B > 0x7fffffffLL - 1024) return fail_msg(IBLDPC_E_INVALID, "bad B");
    ECK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int wp = (int)(((B + 31) / 32 + 3) / 4 * 4);   // words per row, padded to a uint4
    const size_t need = (size_t)wp * ((size_t)h->K + 2 * (size_t)h->M);
    if (h->scratch_words < need) {
        if (h->d_x) ECK(cudaFree(h->d_x));
        h->d_x = nullptr;
        h->scratch_words = 0;
        if (cudaMalloc((void**)&h->d_x, need * sizeof(uint32_t)) != cudaSuccess)
            return fail_msg(IBLDPC_E_NOMEM, "cudaMalloc of the encoder scratch failed");
        h->scratch_words = need;
    }
    h->d_s = h->d_x + (size_t)wp * h->K;
    h->d_p = h->d_s + (size_t)wp * h->M;
    {
        const long long warps = (long long)h->K * wp;
        const int grid = (int)std::min<long long>((warps * 32 + 255) / 256, 148LL * 32);
        pack_bits_kernel<<<grid, 256, 0, st>>>(bits_dev, h->d_x, h->K, B, wp);
    }
    {
        dim3 grid((wp + 127) / 128, (unsigned)std::min(h->M, 16384));
        gf2_spmv_kernel<<<grid, 128, 0, st>>>(h->d_a_rowptr, h->d_a_col, h->d_x, h->d_s, h->M, wp);
    }
    if (h->method == IBLDPC_ENC_SUBSTITUTION) {
        gf2_trisolve_kernel<<<(wp + 31) / 32, 32, 0, st>>>(h->d_eq, h->d_var, h->d_optr, h->d_oth, h->d_s, h->d_p, h->M, wp);
    } else {
        dim3 grid((wp / 4 + 63) / 64, (unsigned)std::min(h->M, 16384));
        gf2_dense_kernel<<<grid, 64, 0, st>>>(h->d_G, h->gw, h->d_s, h->d_p, h->M, wp / 4);
    }
    {
        const long long n = (long long)h->N * B;
        const int grid = (int)std::min<long long>((n + 255) / 256, 148LL * 32);
        unpack_codeword_kernel<<<grid, 256, 0, st>>>(bits_dev, h->d_p, codeword_dev, h->K, h->N, B, wp);
    }
    ECK(cudaGetLastError());
    return IBLDPC_OK;
}

int ibldpc_encoder_destroy(ibldpc_encoder_handle h)
{
    if (!h) return IBLDPC_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (int* p : {h->d_a_rowptr, h->d_a_col, h->d_eq, h->d_var, h->d_optr, h->d_oth})
        if (p) cudaFree(p);
    if (h->d_G) cudaFree(h->d_G);
    if (h->d_x) cudaFree(h->d_x);
    delete h;
    return IBLDPC_OK;
}

int ibldpc_random_bits(int device, uint64_t seed, uint64_t offset, int64_t n, uint8_t* out_dev, void* stream)
{
    if (n < 0 || (n > 0 && !out_dev)) return fail_msg(IBLDPC_E_INVALID, "bad arguments");
    if (n == 0) return IBLDPC_OK;
    ECK(cudaSetDevice(device));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    channel_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(nullptr, n, 0.0, seed, offset, out_dev);
    ECK(cudaGetLastError());
    return IBLDPC_OK;
}

int ibldpc_awgn(int device, const double* x_dev, const uint8_t* bits_dev, int64_t n, double sigma_n2, uint64_t seed,
                uint64_t offset, double* y_dev, void* stream)
{
    if (n < 0 || (n > 0 && (!y_dev || (!x_dev == !bits_dev))) || !(sigma_n2 >= 0.0))
        return fail_msg(IBLDPC_E_INVALID, "bad arguments (exactly one of x_dev / bits_dev, sigma_n2 >= 0)");
    if (n == 0) return IBLDPC_OK;
    ECK(cudaSetDevice(device));
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
    const double sigma = sqrt(sigma_n2);
    if (x_dev) channel_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, sigma, seed, offset, y_dev);
    else channel_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(bits_dev, n, sigma, seed, offset, y_dev);
    ECK(cudaGetLastError());
    return IBLDPC_OK;
}

}  // extern "C"
