// llr_layered.cu -- opt-in LAYERED (row-message-passing) schedule for the min-sum / belief-propagation benchmark decoders
// ("layered/row-schedule variants ... opt-in because they change results vs the reference").
//
// The reference decoders run the flooding schedule only (min_sum_decoder_irreg.py:242-273: every check node, then every
// variable node).  A layered decoder keeps ONE a-posteriori LLR per variable node and updates it check by check, so a
// check already sees the messages the checks before it produced in the same iteration; it needs about half the
// iterations of flooding for the same error rate.  Semantics here (there is no reference implementation to match; the
// numpy restatement in tests/test_gpu_round2.py is the checker, and the node arithmetic is the reference's):
//   L[v] = channel LLR, R[e] = 0
//   pass it = 0 .. imax-2 (at most imax-1 passes, like the reference loop):
//     for every LAYER (a set of checks that share no variable; greedy colouring in check order, fixed per code):
//       for every check c of the layer, edges k = 0..d-1 in slot order (ascending variable index):
//         X_k = L[v_k] - R[e_k]                              the variable-to-check message, exact
//         R[e_k] = checknode(clip150(X_0..X_{d-1}) without k)  min-sum (:126-167) or box-plus (:32-71, forward/backward);
//                                                            the check sees the message clipped like the reference's
//                                                            variable-node output (kernels_min_and_BP.cl:120)
//         L[v_k] = X_k + R[e_k]                              so L = channel + sum of the current R at all times, unclipped
//                                                            like the reference's output; clipping X itself would drop
//                                                            the other checks' contributions from L and the recursion
//                                                            oscillates once the messages saturate
//     syndrome of the hard decisions (L < 0) of all checks; early termination stops the batch when it is zero
//   output: L (unclipped a-posteriori LLRs, like calc_varnode_output :170-204); i_num = passes executed + 1.
// Within a layer the checks are independent, so a layer is one launch per degree class: one warp = one (check, tile of
// 32 x VEC frames), frames across lanes, 128-bit accesses -- the mapping of llr_cn_kernel.  Layers are visited in colour
// order, checks inside a layer in any order (they touch disjoint variables): results do not depend on the launch geometry.
#include <algorithm>
#include <string>
#include <vector>

#include "ibldpc_internal.h"
#include "kernel_tables.h"

namespace ibldpc {

struct LayerPlan {
    int n_layers = 0;
    struct Group { int layer, degree, count; int* d_nodes; };
    std::vector<Group> groups;          // in launch order: layer by layer, degree classes inside a layer
    int* d_all = nullptr;               // one allocation behind all node lists
};

void layered_free(ibldpc_decoder* h)
{
    if (!h->layers) return;
    if (h->layers->d_all) cudaFree(h->layers->d_all);
    delete h->layers;
    h->layers = nullptr;
}

namespace {

template <typename F, int ALGO, int D>
__global__ void __launch_bounds__(kThreads) llr_layer_kernel(LlrArgs a, const int* __restrict__ nodes, int n_nodes)
{
    if (a.early && a.it >= 1 && a.flags[a.it - 1] == 0) return;   // batch already converged
    constexpr int V = VecOf<F>::N;
    constexpr int MAXD = D > 0 ? D : kMaxGenericDeg;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const long long items = (long long)n_nodes * a.tiles;
    F* R = static_cast<F*>(a.cin);     // check-node-major check-to-variable messages
    F* L = static_cast<F*>(a.out);     // a-posteriori LLRs
    for (long long item = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5); item < items; item += nwarps) {
        const int i = (int)(item / a.tiles);
        const int tile = (int)(item - (long long)i * a.tiles);
        const long long col = ((long long)tile * 32 + lane) * V;
        if (col >= a.pitch) continue;
        const int c = nodes[i];
        const int s = a.sc[c];
        const int d = D > 0 ? D : a.deg_c[c];
        Vec<F> x[MAXD], q[MAXD], o[MAXD];   // x = L - R_old exactly, q = what the check node sees
        int v[MAXD];
#pragma unroll
        for (int k = 0; k < (D > 0 ? D : d); ++k) {
            v[k] = a.vidx[s + k];
            const Vec<F> l = vload(L + (long long)v[k] * a.pitch + col);
            const Vec<F> r = vload(R + (long long)(s + k) * a.pitch + col);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                x[k].v[e] = l.v[e] - r.v[e];
                q[k].v[e] = clip150(x[k].v[e]);
            }
        }
        llr_cn_compute<F, ALGO, D>(q, o, d);
#pragma unroll
        for (int k = 0; k < (D > 0 ? D : d); ++k) {
            vstore(R + (long long)(s + k) * a.pitch + col, o[k]);
            Vec<F> l;
#pragma unroll
            for (int e = 0; e < V; ++e) l.v[e] = x[k].v[e] + o[k].v[e];
            vstore(L + (long long)v[k] * a.pitch + col, l);
        }
    }
}

// syndrome of the hard decisions of the a-posteriori LLRs: any unsatisfied check of a valid frame raises flags[it]
template <typename F>
__global__ void __launch_bounds__(kThreads) llr_app_syndrome_kernel(LlrArgs a)
{
    if (a.it >= 1 && a.flags[a.it - 1] == 0) return;
    constexpr int V = VecOf<F>::N;
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * kWarpsPerCta;
    const long long items = (long long)a.n_chk * a.tiles;
    const F* L = static_cast<const F*>(a.out);
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
            const Vec<F> l = vload(L + (long long)a.vidx[s + k] * a.pitch + col);
#pragma unroll
            for (int e = 0; e < V; ++e) par[e] ^= (l.v[e] < F(0)) ? 1 : 0;
        }
#pragma unroll
        for (int e = 0; e < V; ++e) bad |= par[e] && (col + e < a.B);
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&a.flags[a.it], 1);
}

__global__ void llr_layered_inum_kernel(const int* flags, int imax, int early, int* inum)
{
    int passes = imax - 1;
    if (early)
        for (int it = 0; it < imax - 1; ++it)
            if (flags[it] == 0) { passes = it + 1; break; }
    *inum = passes + 1;
}

template <typename F, int ALGO>
LlrNodeKernel layer_kernel_for(int d)
{
    switch (d) {
    case 2: return llr_layer_kernel<F, ALGO, 2>;
    case 3: return llr_layer_kernel<F, ALGO, 3>;
    case 4: return llr_layer_kernel<F, ALGO, 4>;
    case 5: return llr_layer_kernel<F, ALGO, 5>;
    case 6: return llr_layer_kernel<F, ALGO, 6>;
    case 7: return llr_layer_kernel<F, ALGO, 7>;
    case 8: return llr_layer_kernel<F, ALGO, 8>;
    case 9: return llr_layer_kernel<F, ALGO, 9>;
    case 10: return llr_layer_kernel<F, ALGO, 10>;
    default: return llr_layer_kernel<F, ALGO, 0>;
    }
}

// Greedy colouring of the checks in index order: a check takes the lowest colour none of the checks it shares a variable
// with has taken.  Deterministic for a given code; quasi-cyclic codes get (close to) their block rows back.
int build_layers(ibldpc_decoder* h)
{
    const int M = h->M, N = h->N;
    const std::vector<int>& sc = h->h_sc;
    const std::vector<int>& vidx = h->h_vidx;
    const std::vector<int>& dc = h->h_dc;
    // checks of every variable
    std::vector<int> vstart((size_t)N + 1, 0);
    for (int e = 0; e < h->E; ++e) vstart[vidx[e] + 1]++;
    for (int v = 0; v < N; ++v) vstart[v + 1] += vstart[v];
    std::vector<int> vchk((size_t)h->E), fill(vstart.begin(), vstart.end() - 1);
    for (int c = 0; c < M; ++c)
        for (int k = 0; k < dc[c]; ++k) vchk[fill[vidx[sc[c] + k]]++] = c;
    std::vector<int> colour((size_t)M, -1), mark;
    int n_layers = 0;
    for (int c = 0; c < M; ++c) {
        mark.assign((size_t)n_layers + 1, 0);
        for (int k = 0; k < dc[c]; ++k) {
            const int v = vidx[sc[c] + k];
            for (int j = vstart[v]; j < vstart[v + 1]; ++j) {
                const int o = colour[vchk[j]];
                if (o >= 0) mark[o] = 1;
            }
        }
        int col = 0;
        while (col < n_layers && mark[col]) ++col;
        colour[c] = col;
        n_layers = std::max(n_layers, col + 1);
    }
    LayerPlan* p = new LayerPlan();
    p->n_layers = n_layers;
    std::vector<int> all;
    all.reserve((size_t)M);
    struct Tmp { int layer, degree, offset, count; };
    std::vector<Tmp> tmp;
    for (int l = 0; l < n_layers; ++l) {
        std::vector<int> degs;
        for (int c = 0; c < M; ++c)
            if (colour[c] == l && std::find(degs.begin(), degs.end(), dc[c]) == degs.end()) degs.push_back(dc[c]);
        std::sort(degs.rbegin(), degs.rend());
        for (int d : degs) {
            Tmp t{l, d, (int)all.size(), 0};
            for (int c = 0; c < M; ++c)
                if (colour[c] == l && dc[c] == d) all.push_back(c);
            t.count = (int)all.size() - t.offset;
            tmp.push_back(t);
        }
    }
    if (cudaMalloc((void**)&p->d_all, sizeof(int) * (size_t)std::max(M, 1)) != cudaSuccess) {
        delete p;
        return fail_msg(IBLDPC_E_NOMEM, "cudaMalloc of the layer lists failed");
    }
    if (cudaMemcpy(p->d_all, all.data(), sizeof(int) * (size_t)M, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(p->d_all);
        delete p;
        return fail_msg(IBLDPC_E_CUDA, "upload of the layer lists failed");
    }
    for (const Tmp& t : tmp) p->groups.push_back(LayerPlan::Group{t.layer, t.degree, t.count, p->d_all + t.offset});
    h->layers = p;
    return IBLDPC_OK;
}

template <typename F, int ALGO>
int decode_layered_padded(ibldpc_decoder* h, Workspace& w, const F* ch, long long pitch, long long B, int imax, int early, F* out,
                          cudaStream_t st)
{
    constexpr int V = VecOf<F>::N;
    const size_t need = (size_t)h->E * (size_t)pitch * sizeof(F);
    if (w.llr_bytes < need || !w.cin) {
        if (w.cin) IBLDPC_CK(cudaFree(w.cin));
        if (w.vin) IBLDPC_CK(cudaFree(w.vin));
        w.cin = w.vin = nullptr;
        w.llr_bytes = 0;
        if (cudaMalloc(&w.cin, need) != cudaSuccess || cudaMalloc(&w.vin, need) != cudaSuccess)
            return fail_msg(IBLDPC_E_NOMEM, "cudaMalloc of LLR message arrays failed");
        w.llr_bytes = need;
    }
    IBLDPC_CK(cudaMemsetAsync(w.flags + 1, 0, sizeof(int) * (size_t)std::max(imax, 1), st));
    IBLDPC_CK(cudaMemsetAsync(w.cin, 0, need, st));                                             // R = 0
    if (out != ch) IBLDPC_CK(cudaMemcpyAsync(out, ch, (size_t)h->N * pitch * sizeof(F), cudaMemcpyDeviceToDevice, st));   // L = channel
    LlrArgs a{};
    a.sc = h->d_sc; a.deg_c = h->d_dc; a.tc = h->d_tc; a.sv = h->d_sv; a.deg_v = h->d_dv; a.tv = h->d_tv; a.vidx = h->d_vidx;
    a.n_var = h->N; a.n_chk = h->M;
    a.ch = ch; a.cin = w.cin; a.vin = w.vin; a.out = out;
    a.pitch = pitch; a.B = (int)B; a.tiles = (int)((pitch + 32 * V - 1) / (32 * V));
    a.flags = w.flags + 1; a.inum = w.inum; a.early = early; a.imax = imax;
    h->last_launches = 0;
    auto grid_of = [&](long long warps) { return (int)std::max<long long>(1, std::min<long long>((warps + kWarpsPerCta - 1) / kWarpsPerCta, (long long)h->sm_count * 8)); };
    for (int it = 0; it < imax - 1; ++it) {
        LlrArgs b = a;
        b.it = it;
        for (const LayerPlan::Group& g : h->layers->groups) {
            LlrNodeKernel k = layer_kernel_for<F, ALGO>(g.degree);
            k<<<grid_of((long long)g.count * b.tiles), kThreads, 0, st>>>(b, g.d_nodes, g.count);
            h->last_launches++;
        }
        if (early) {
            llr_app_syndrome_kernel<F><<<grid_of((long long)h->M * b.tiles), kThreads, 0, st>>>(b);
            h->last_launches++;
        }
    }
    llr_layered_inum_kernel<<<1, 1, 0, st>>>(a.flags, imax, early, a.inum);
    h->last_launches++;
    IBLDPC_CK(cudaGetLastError());
    return IBLDPC_OK;
}

}  // namespace

int layered_prepare(ibldpc_decoder* h)
{
    if (h->layers) return IBLDPC_OK;
    if (h->dc_max > kMaxGenericDeg) return fail_msg(IBLDPC_E_INVALID, "check-node degree too large for the layered kernels");
    return build_layers(h);
}

int layered_count(const ibldpc_decoder* h) { return h->layers ? h->layers->n_layers : 0; }

// algo: IBLDPC_ALGO_MINSUM or IBLDPC_ALGO_BP (forward/backward box-plus, the float64 default of the flooding decoder)
int decode_llr_layered_f32(ibldpc_decoder* h, Workspace& w, int algo, const float* ch, long long pitch, long long B, int imax,
                           int early, float* out, cudaStream_t st)
{
    return algo == IBLDPC_ALGO_MINSUM ? decode_layered_padded<float, 0>(h, w, ch, pitch, B, imax, early, out, st)
                                      : decode_layered_padded<float, 2>(h, w, ch, pitch, B, imax, early, out, st);
}
int decode_llr_layered_f64(ibldpc_decoder* h, Workspace& w, int algo, const double* ch, long long pitch, long long B, int imax,
                           int early, double* out, cudaStream_t st)
{
    return algo == IBLDPC_ALGO_MINSUM ? decode_layered_padded<double, 0>(h, w, ch, pitch, B, imax, early, out, st)
                                      : decode_layered_padded<double, 2>(h, w, ch, pitch, B, imax, early, out, st);
}

}  // namespace ibldpc
