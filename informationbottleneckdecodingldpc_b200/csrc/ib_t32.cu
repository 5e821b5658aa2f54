// ib_t32.cu -- host side of the |T| <= 32 shared-memory family (ib_kernels_t32.cuh): table images of every
// (iteration, degree class), built once at ibldpc_set_luts, and the launch sequence of one decode.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ibldpc_internal.h"
#include "ib_kernels_t32.cuh"

namespace ibldpc {

struct T32Images {
    uint8_t* d_images = nullptr;
    // byte offsets inside d_images: [class][iteration]
    std::vector<std::vector<size_t>> cn_off, vn_off, out_off;
    std::vector<int*> cn_starts, vn_starts;
    bool attrs_set = false;
    bool phase_attrs_set = false;
    bool coop_attrs_set = false;
};

namespace {

// image of one class: [striped stages (the last ones)][plain stages (the first ones)], see ib_kernels_t32.cuh
void build_t32_image(uint8_t* img, int nst, int T, const uint8_t* stages /* [nst][T*T] */, size_t stage_stride,
                     const uint8_t* match_row)
{
    memset(img, 0, (size_t)t32_image_bytes(nst));
    for (int c = 0; c < nst; ++c) {
        const uint8_t* S = stages + (size_t)c * stage_stride;
        uint8_t* dst = img + t32_col_base(nst, c);
        const bool fold = match_row != nullptr && c == nst - 1;
        for (int t = 0; t < T; ++t)
            for (int m = 0; m < T; ++m) {
                uint8_t e = S[t * T + m];
                if (fold) e = match_row[e];
                if (t32_is_striped(nst, c)) {
                    uint8_t* row = dst + (size_t)t * 1024 + (size_t)(m >> 2) * 128 + (m & 3);
                    for (int l = 0; l < 32; ++l) row[l * 4] = e;
                } else {
                    dst[t * 32 + m] = e;
                }
            }
    }
}

}  // namespace

void t32_free(ibldpc_decoder* h)
{
    T32Images* p = h->t32_images;
    if (!p) return;
    if (p->d_images) cudaFree(p->d_images);
    for (int* q : p->cn_starts) if (q) cudaFree(q);
    for (int* q : p->vn_starts) if (q) cudaFree(q);
    delete p;
    h->t32_images = nullptr;
}

int t32_prepare(ibldpc_decoder* h)
{
    t32_free(h);
    if (!h->t32) return IBLDPC_OK;
    T32Images* p = new T32Images();
    h->t32_images = p;
    const int T = h->T, TT = T * T, DC = h->DC, DV = h->DV, imax = h->lut_imax;
    std::vector<uint8_t> host;
    auto add = [&](int nst, const uint8_t* stages, const uint8_t* mrow) -> size_t {
        const size_t off = host.size();
        host.resize(off + (size_t)t32_image_bytes(nst));
        if (nst > 0) build_t32_image(host.data() + off, nst, T, stages, (size_t)TT, mrow);
        return off;
    };
    for (auto& c : h->cn_classes) {
        std::vector<size_t> offs;
        for (int blk = 0; blk < imax; ++blk)
            offs.push_back(add(c.degree - 2, h->h_cn8.data() + (size_t)blk * (DC - 2) * TT,
                               h->match ? h->h_mc8.data() + ((size_t)blk * DC + (c.degree - 1)) * T : nullptr));
        p->cn_off.push_back(offs);
    }
    for (auto& c : h->vn_classes) {
        std::vector<size_t> up, dec;
        for (int it = 0; it < imax; ++it) {
            up.push_back(add(c.degree - 1, h->h_vn8.data() + (size_t)it * DV * TT,
                             h->match ? h->h_mv8.data() + ((size_t)it * DV + (c.degree - 1)) * T : nullptr));
        }
        for (int it = 0; it < imax; ++it) dec.push_back(add(c.degree, h->h_vn8.data() + (size_t)it * DV * TT, nullptr));
        p->vn_off.push_back(up);
        p->out_off.push_back(dec);
    }
    IBLDPC_CK(cudaMalloc((void**)&p->d_images, std::max<size_t>(host.size(), 16)));
    IBLDPC_CK(cudaMemcpy(p->d_images, host.data(), host.size(), cudaMemcpyHostToDevice));
    auto starts_of = [&](const NodeClass& c, const std::vector<int>& all_starts, std::vector<int*>& out) -> int {
        std::vector<int> nodes((size_t)c.count), st((size_t)c.count);
        IBLDPC_CK(cudaMemcpy(nodes.data(), c.d_nodes, sizeof(int) * (size_t)c.count, cudaMemcpyDeviceToHost));
        for (int i = 0; i < c.count; ++i) st[i] = all_starts[nodes[i]];
        int* d = nullptr;
        IBLDPC_CK(cudaMalloc((void**)&d, sizeof(int) * (size_t)std::max(c.count, 1)));
        IBLDPC_CK(cudaMemcpy(d, st.data(), sizeof(int) * (size_t)c.count, cudaMemcpyHostToDevice));
        out.push_back(d);
        return IBLDPC_OK;
    };
    for (auto& c : h->cn_classes)
        if (int rc = starts_of(c, h->h_sc, p->cn_starts)) return rc;
    for (auto& c : h->vn_classes)
        if (int rc = starts_of(c, h->h_sv, p->vn_starts)) return rc;
    return IBLDPC_OK;
}

// One decode; `a` = graph pointers, sanitised uint8 channel buffer, message array, output, pitch, flags.
int decode_ib_t32(ibldpc_decoder* h, const IbArgs& a, int imax, int early, cudaStream_t st)
{
    T32Images* p = h->t32_images;
    if (!p->attrs_set) {
        for (auto& c : h->cn_classes)
            for (int e = 0; e < 2; ++e)
                IBLDPC_CK(cudaFuncSetAttribute((const void*)t32_cn_kernel(c.degree, e != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               t32_image_bytes(c.degree - 2)));
        for (auto& c : h->vn_classes) {
            IBLDPC_CK(cudaFuncSetAttribute((const void*)t32_vn_kernel(c.degree), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           t32_image_bytes(c.degree - 1)));
            IBLDPC_CK(cudaFuncSetAttribute((const void*)t32_out_kernel(c.degree), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           t32_image_bytes(c.degree)));
        }
        p->attrs_set = true;
    }
    auto grid_for = [&](int mode, const NodeClass& c) {
        const long long vec = t32_vec(mode, c.degree);
        const long long tiles = ((long long)a.pitch + 128 * vec - 1) / (128 * vec);
        return (int)std::max<long long>(1, std::min<long long>(h->sm_count, ((long long)c.count * tiles + 31) / 32));
    };
    auto prof_begin = [&](int phase) -> int {
        if (!h->profiling) return IBLDPC_OK;
        PhaseEvent ev;
        ev.phase = phase;
        IBLDPC_CK(cudaEventCreate(&ev.a));
        IBLDPC_CK(cudaEventCreate(&ev.b));
        IBLDPC_CK(cudaEventRecord(ev.a, st));
        h->events.push_back(ev);
        return IBLDPC_OK;
    };
    auto prof_end = [&]() -> int {
        if (h->profiling) IBLDPC_CK(cudaEventRecord(h->events.back().b, st));
        return IBLDPC_OK;
    };
    // ---- one launch per phase over all degree classes (ib_t32_phase_kernel) for degree sets of up to four classes
    const bool no_fused = getenv("IBLDPC_T32_NO_PHASE") != nullptr;   // one launch per degree class (A/B, parity tests)
    if (!no_fused && h->cn_classes.size() <= (size_t)kT32MaxClasses && h->vn_classes.size() <= (size_t)kT32MaxClasses) {
        int cn_smem = 16, vn_smem = 16, out_smem = 16;
        for (auto& c : h->cn_classes) cn_smem = std::max(cn_smem, t32_image_bytes(c.degree - 2));
        for (auto& c : h->vn_classes) {
            vn_smem = std::max(vn_smem, c.degree > 1 ? t32_image_bytes(c.degree - 1) : 16);
            out_smem = std::max(out_smem, t32_image_bytes(c.degree));
        }
        if (!p->phase_attrs_set) {
            for (int e = 0; e < 2; ++e)
                IBLDPC_CK(cudaFuncSetAttribute((const void*)t32_phase_cn_kernel(e != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, cn_smem));
            IBLDPC_CK(cudaFuncSetAttribute((const void*)t32_phase_vn_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, vn_smem));
            IBLDPC_CK(cudaFuncSetAttribute((const void*)t32_phase_out_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, out_smem));
            p->phase_attrs_set = true;
        }
        // classes heaviest first
        auto order_of = [](const std::vector<NodeClass>& cls) {
            std::vector<int> o(cls.size());
            for (size_t i = 0; i < cls.size(); ++i) o[i] = (int)i;
            std::sort(o.begin(), o.end(), [&](int x, int y) { return cls[x].degree > cls[y].degree; });
            return o;
        };
        const std::vector<int> cn_order = order_of(h->cn_classes), vn_order = order_of(h->vn_classes);
        auto fill = [&](T32PhaseArgs& q, int mode, const std::vector<NodeClass>& cls, const std::vector<int>& order,
                        const std::vector<int*>& starts) -> int {
            long long chunks = 0;
            q.n_cls = (int)order.size();
            for (size_t k = 0; k < order.size(); ++k) {
                const NodeClass& c = cls[order[k]];
                q.deg[k] = c.degree; q.nodes[k] = c.d_nodes; q.starts[k] = starts[order[k]]; q.n_nodes[k] = c.count;
                const long long vec = t32_vec(mode, c.degree);
                const long long tiles = ((long long)a.pitch + 128 * vec - 1) / (128 * vec);
                chunks += ((long long)c.count * tiles + 31) / 32;
            }
            return (int)std::max<long long>(1, std::min<long long>(h->sm_count, chunks));
        };
        T32PhaseArgs pb{};
        pb.a = a;
        pb.a.early = early;
        pb.a.imax = imax;
        int rc;
        // ---- small batches: the whole decode in one cooperative launch (ib_t32_coop_kernel).  Measured on B200, 802.11n
        // |T| = 32: B = 2000 1.55 -> see profiles/README.md; the per-phase launches win once a phase is long enough
        // (threshold IBLDPC_T32_COOP_MAX_B, default 2048 frames and at most 32 MB of messages)
        const long long coop_max = getenv("IBLDPC_T32_COOP_MAX_B") ? atoll(getenv("IBLDPC_T32_COOP_MAX_B")) : 2048;
        if ((long long)a.B <= coop_max && (long long)h->E * a.pitch <= (32LL << 20)) {
            if (h->coop_supported < 0) {
                int v = 0;
                IBLDPC_CK(cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->device));
                h->coop_supported = v;
            }
            if (h->coop_supported) {
                const int smem = std::max(cn_smem, std::max(vn_smem, out_smem));
                T32CoopKernel k = t32_coop_kernel(early != 0);
                if (!p->coop_attrs_set) {
                    for (int e = 0; e < 2; ++e)
                        IBLDPC_CK(cudaFuncSetAttribute((const void*)t32_coop_kernel(e != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                    p->coop_attrs_set = true;
                }
                T32CoopArgs q{};
                q.a = pb.a;
                const int g_cn = fill(q.cn, kPhaseCn, h->cn_classes, cn_order, p->cn_starts);
                const int g_vn = fill(q.vn, kPhaseVn, h->vn_classes, vn_order, p->vn_starts);
                const int g_out = fill(q.out, kPhaseOut, h->vn_classes, vn_order, p->vn_starts);
                const long long one = h->lut_imax > 1 ? 1 : 0;
                for (size_t kk = 0; kk < cn_order.size(); ++kk) {
                    const int ci = cn_order[kk];
                    q.cn.image[kk] = p->d_images + p->cn_off[ci][0];
                    q.cn.image_stride[kk] = one ? (long long)(p->cn_off[ci][1] - p->cn_off[ci][0]) : 0;
                }
                for (size_t kk = 0; kk < vn_order.size(); ++kk) {
                    const int ci = vn_order[kk];
                    q.vn.image[kk] = p->d_images + p->vn_off[ci][0];
                    q.vn.image_stride[kk] = one ? (long long)(p->vn_off[ci][1] - p->vn_off[ci][0]) : 0;
                    q.out.image[kk] = p->d_images + p->out_off[ci][0];
                    q.out.image_stride[kk] = one ? (long long)(p->out_off[ci][1] - p->out_off[ci][0]) : 0;
                }
                const int grid = std::max(g_cn, std::max(g_vn, g_out));
                if ((rc = prof_begin(2))) return rc;
                void* params[] = {&q};
                IBLDPC_CK(cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(kT32Threads), params, (size_t)smem, st));
                h->last_launches++; h->last_grid = grid; h->last_smem = smem;
                if ((rc = prof_end())) return rc;
                return IBLDPC_OK;
            }
        }
        auto phase_cn = [&](int it) -> int {
            if ((rc = prof_begin(it < 0 ? 2 : 0))) return rc;
            T32PhaseArgs q = pb;
            q.a.it = it;
            q.a.iter0 = it < 0;
            const int grid = fill(q, kPhaseCn, h->cn_classes, cn_order, p->cn_starts);
            for (size_t k = 0; k < cn_order.size(); ++k) q.image[k] = p->d_images + p->cn_off[cn_order[k]][it + 1];
            t32_phase_cn_kernel(early != 0)<<<grid, kT32Threads, cn_smem, st>>>(q);
            h->last_launches++; h->last_grid = grid; h->last_smem = cn_smem;
            return prof_end();
        };
        auto phase_vn = [&](int it, bool decide) -> int {
            if ((rc = prof_begin(decide ? 2 : 1))) return rc;
            T32PhaseArgs q = pb;
            q.a.it = it;
            q.a.iter0 = 0;
            const int grid = fill(q, decide ? kPhaseOut : kPhaseVn, h->vn_classes, vn_order, p->vn_starts);
            for (size_t k = 0; k < vn_order.size(); ++k) {
                const int ci = vn_order[k];
                if (decide) {
                    q.image[k] = p->d_images + p->out_off[ci][0];
                    q.image_stride[k] = h->lut_imax > 1 ? (long long)(p->out_off[ci][1] - p->out_off[ci][0]) : 0;
                } else {
                    q.image[k] = p->d_images + p->vn_off[ci][it];
                }
            }
            if (decide) t32_phase_out_kernel()<<<grid, kT32Threads, out_smem, st>>>(q);
            else t32_phase_vn_kernel()<<<grid, kT32Threads, vn_smem, st>>>(q);
            h->last_launches++;
            return prof_end();
        };
        if ((rc = phase_cn(-1))) return rc;
        for (int it = 0; it < imax - 1; ++it) {
            if ((rc = phase_vn(it, false))) return rc;
            if ((rc = phase_cn(it))) return rc;
        }
        if ((rc = phase_vn(0, true))) return rc;
        IBLDPC_CK(cudaGetLastError());
        return IBLDPC_OK;
    }
    T32Args base{};
    base.a = a;
    base.a.early = early;
    base.a.imax = imax;
    int rc;
    auto launch_cn = [&](int it) -> int {
        if ((rc = prof_begin(it < 0 ? 2 : 0))) return rc;
        for (size_t ci = 0; ci < h->cn_classes.size(); ++ci) {
            const NodeClass& c = h->cn_classes[ci];
            T32Args q = base;
            q.a.it = it;
            q.a.iter0 = it < 0;
            q.image = p->d_images + p->cn_off[ci][it + 1];
            q.nodes = c.d_nodes; q.starts = p->cn_starts[ci]; q.n_nodes = c.count;
            const int grid = grid_for(kPhaseCn, c);
            t32_cn_kernel(c.degree, early != 0)<<<grid, kT32Threads, t32_image_bytes(c.degree - 2), st>>>(q);
            h->last_launches++; h->last_grid = grid; h->last_smem = t32_image_bytes(c.degree - 2);
        }
        return prof_end();
    };
    auto launch_vn = [&](int it, bool decide) -> int {
        if ((rc = prof_begin(decide ? 2 : 1))) return rc;
        for (size_t ci = 0; ci < h->vn_classes.size(); ++ci) {
            const NodeClass& c = h->vn_classes[ci];
            T32Args q = base;
            q.a.it = it;
            q.a.iter0 = 0;
            q.nodes = c.d_nodes; q.starts = p->vn_starts[ci]; q.n_nodes = c.count;
            if (decide) {
                q.image = p->d_images + p->out_off[ci][0];
                q.image_stride = h->lut_imax > 1 ? (long long)(p->out_off[ci][1] - p->out_off[ci][0]) : 0;
                t32_out_kernel(c.degree)<<<grid_for(kPhaseOut, c), kT32Threads, t32_image_bytes(c.degree), st>>>(q);
            } else {
                q.image = p->d_images + p->vn_off[ci][it];
                const int smem = c.degree > 1 ? t32_image_bytes(c.degree - 1) : 0;
                t32_vn_kernel(c.degree)<<<grid_for(kPhaseVn, c), kT32Threads, smem, st>>>(q);
            }
            h->last_launches++;
        }
        return prof_end();
    };
    if ((rc = launch_cn(-1))) return rc;
    for (int it = 0; it < imax - 1; ++it) {
        if ((rc = launch_vn(it, false))) return rc;
        if ((rc = launch_cn(it))) return rc;
    }
    if ((rc = launch_vn(0, true))) return rc;
    IBLDPC_CK(cudaGetLastError());
    return IBLDPC_OK;
}

}  // namespace ibldpc
