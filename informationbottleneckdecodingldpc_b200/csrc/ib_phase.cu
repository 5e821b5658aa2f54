// ib_phase.cu -- host side of the fused per-phase kernels (ib_phase_n4.cuh): pre-expanded shared-memory images of
// every phase, built once at ibldpc_set_luts, and the launch sequence of one decode (one launch per phase).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "ibldpc_internal.h"
#include "ib_phase_sets.h"

namespace ibldpc {

struct PhaseImages {
    const PhaseSetOps* ops = nullptr;
    int imax = 0;
    uint8_t* d_images = nullptr;            // [imax CN blocks][imax VN update images][imax decision images]
    size_t cn_bytes = 0, vn_bytes = 0, out_bytes = 0;
    // node lists in the order of the degree set (heaviest first) and their inbox start offsets
    const int* cn_nodes[kPhaseMaxClasses] = {};
    const int* vn_nodes[kPhaseMaxClasses] = {};
    int* cn_starts[kPhaseMaxClasses] = {};
    int* vn_starts[kPhaseMaxClasses] = {};
    int cn_count[kPhaseMaxClasses] = {}, vn_count[kPhaseMaxClasses] = {};
    int occ_checked = 0;
    int coop_checked = 0;
};

namespace {

bool same_degrees(const std::vector<int>& set, const std::vector<NodeClass>& classes)
{
    if (set.size() != classes.size()) return false;
    for (int d : set) {
        bool found = false;
        for (auto& c : classes) found |= c.degree == d;
        if (!found) return false;
    }
    return true;
}

int class_index(const std::vector<NodeClass>& classes, int degree)
{
    for (size_t i = 0; i < classes.size(); ++i)
        if (classes[i].degree == degree) return (int)i;
    return -1;
}

// One image: [pair regions][256 rows (m*16+t) x words x 32 lanes x 4 stage columns].
//   stage(j)      -> T*T bytes of look-up stage j of this phase (reference order t*T+m)
//   match_row(d)  -> T bytes of the matching row of degree d, or nullptr
//   pair_rows(ci) -> T*T x 8 bytes of the composed tail-pair rows of class index ci (handle order), or nullptr
//   tri_table     -> 16^3 bytes of the three-input table of this phase (ib_triple_n4.cuh), classes written d + 300 only
template <typename StageFn, typename MatchFn, typename PairFn>
void build_image(uint8_t* img, const PhaseLayoutRt& L, int mode, int T, const std::vector<NodeClass>& classes, StageFn stage,
                 MatchFn match_row, PairFn pair_rows, const uint8_t* tri_table = nullptr)
{
    memset(img, 0, (size_t)L.image_bytes);
    uint32_t* tab = reinterpret_cast<uint32_t*>(img + (size_t)L.tab_offset);
    if (L.tab_offset != L.tri_offset && tri_table != nullptr) {
        // per-lane replicated: word q of the shared-memory table = word q >> 5 of the compact one
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tri_table);
        uint32_t* dst = reinterpret_cast<uint32_t*>(img + (size_t)L.tri_offset);
        for (int q = 0; q < kTripleBytes / 4; ++q) dst[q] = src[q >> 5];
    }
    const int W = L.words;
    for (int i = 0; i < L.n; ++i) {
        const PhaseClassLayout& c = L.cls[i];
        const int D = c.degree;
        // pair rows: kPairSlots lane slots of 8 bytes per row a*16+b
        auto expand_rows = [&](const uint8_t* src, int region) {
            uint8_t* dst = img + (size_t)region * kPairBytes;
            for (int ra = 0; ra < T; ++ra)
                for (int rb = 0; rb < T; ++rb)
                    for (int s = 0; s < kPairSlots; ++s)
                        memcpy(dst + ((size_t)(ra * kTS + rb) * kPairSlots + s) * 8, src + (size_t)(ra * T + rb) * 8, 8);
        };
        if (c.pair) expand_rows(pair_rows(class_index(classes, D)), c.pair_index);
        // the column stored as 4*x feeds the pair row: check nodes column D-5, variable nodes column D-4 (local index)
        const int x4 = !c.pair ? -1 : (mode == kPhaseCn ? D - 5 : D - 4);
        const uint8_t* mrow = mode == kPhaseOut ? nullptr : match_row(D);
        const int ncols = mode == kPhaseOut ? D : c.cols;
        for (int j = 0; j < ncols; ++j) {
            const uint8_t* S = stage(j);
            const int col = c.col_base + j;
            const int wq = col >> 2, q = col & 3;
            const bool fold = mrow != nullptr && j == ncols - 1;
            for (int m = 0; m < T; ++m)
                for (int t = 0; t < T; ++t) {
                    uint32_t e = S[t * T + m];
                    if (fold) e = mrow[e];
                    if (j == x4) e *= 4u;
#ifdef IBLDPC_DP4A
                    uint32_t* row = tab + ((size_t)(t * kTS + m) * W + wq) * 32;   // rows t*16 + m (dp4a address arithmetic)
#else
                    uint32_t* row = tab + ((size_t)(m * kTS + t) * W + wq) * 32;
#endif
                    const uint32_t mask = 0xffu << (8 * q);
                    for (int l = 0; l < 32; ++l) row[l] = (row[l] & ~mask) | (e << (8 * q));
                }
        }
    }
}

}  // namespace

static void free_images(PhaseImages*& p)
{
    if (!p) return;
    if (p->d_images) cudaFree(p->d_images);
    for (int i = 0; i < kPhaseMaxClasses; ++i) {
        if (p->cn_starts[i]) cudaFree(p->cn_starts[i]);
        if (p->vn_starts[i]) cudaFree(p->vn_starts[i]);
    }
    delete p;
    p = nullptr;
}

void phase_free(ibldpc_decoder* h)
{
    free_images(h->phase);
    free_images(h->phase_tri);
}

// images the per-frame-early-termination kernels run on
static PhaseImages* pf_images(const ibldpc_decoder* h) { return h->phase_tri ? h->phase_tri : h->phase; }

// expands every phase image of the degree set `ops` on the host and uploads them
static int build_phase_images(ibldpc_decoder* h, const PhaseSetOps* ops, PhaseImages** out)
{
    PhaseImages* p = new PhaseImages();
    *out = p;
    p->ops = ops;
    p->imax = h->lut_imax;
    const int T = h->T, TT = T * T, DC = h->DC, DV = h->DV, imax = h->lut_imax;
    p->cn_bytes = (size_t)ops->cn_layout.image_bytes;
    p->vn_bytes = (size_t)ops->vn_layout.image_bytes;
    p->out_bytes = (size_t)ops->out_layout.image_bytes;
    const size_t total = (size_t)imax * (p->cn_bytes + p->vn_bytes + p->out_bytes);
    std::vector<uint8_t> host(total);
    const size_t ncc = h->cn_classes.size(), nvc = h->vn_classes.size();
    for (int blk = 0; blk < imax; ++blk) {
        // check-node phase of table block blk (0 = iteration-0 tables)
        build_image(
            host.data() + (size_t)blk * p->cn_bytes, ops->cn_layout, kPhaseCn, T, h->cn_classes,
            [&](int j) { return h->h_cn8.data() + ((size_t)blk * (DC - 2) + j) * TT; },
            [&](int d) -> const uint8_t* { return h->match ? h->h_mc8.data() + ((size_t)blk * DC + (d - 1)) * T : nullptr; },
            [&](int ci) { return h->h_cn_pair.data() + ((size_t)blk * ncc + ci) * (size_t)TT * 8; },
            h->h_cn3.empty() ? nullptr : h->h_cn3.data() + (size_t)blk * kTripleEntries);
        // variable-node update of iteration blk
        build_image(
            host.data() + (size_t)imax * p->cn_bytes + (size_t)blk * p->vn_bytes, ops->vn_layout, kPhaseVn, T, h->vn_classes,
            [&](int j) { return h->h_vn8.data() + ((size_t)blk * DV + j) * TT; },
            [&](int d) -> const uint8_t* { return h->match ? h->h_mv8.data() + ((size_t)blk * DV + (d - 1)) * T : nullptr; },
            [&](int ci) { return h->h_vn_pair.data() + ((size_t)blk * nvc + ci) * (size_t)TT * 8; },
            h->h_vn3.empty() ? nullptr : h->h_vn3.data() + (size_t)blk * kTripleEntries);
        // decision with the variable-node tables of iteration blk (no message alignment on the output)
        build_image(
            host.data() + (size_t)imax * (p->cn_bytes + p->vn_bytes) + (size_t)blk * p->out_bytes, ops->out_layout, kPhaseOut, T,
            h->vn_classes, [&](int j) { return h->h_vn8.data() + ((size_t)blk * DV + j) * TT; },
            [&](int) -> const uint8_t* { return nullptr; }, [&](int) -> const uint8_t* { return nullptr; });
    }
    IBLDPC_CK(cudaMalloc((void**)&p->d_images, total));
    IBLDPC_CK(cudaMemcpy(p->d_images, host.data(), total, cudaMemcpyHostToDevice));
    // node lists + start offsets in set order
    auto starts_of = [&](const NodeClass& c, const std::vector<int>& all_starts, int** d_out) -> int {
        std::vector<int> nodes((size_t)c.count), st((size_t)c.count);
        IBLDPC_CK(cudaMemcpy(nodes.data(), c.d_nodes, sizeof(int) * (size_t)c.count, cudaMemcpyDeviceToHost));
        for (int i = 0; i < c.count; ++i) st[i] = all_starts[nodes[i]];
        IBLDPC_CK(cudaMalloc((void**)d_out, sizeof(int) * (size_t)std::max(c.count, 1)));
        IBLDPC_CK(cudaMemcpy(*d_out, st.data(), sizeof(int) * (size_t)c.count, cudaMemcpyHostToDevice));
        return IBLDPC_OK;
    };
    for (size_t i = 0; i < ops->cn_deg.size(); ++i) {
        const NodeClass& c = h->cn_classes[class_index(h->cn_classes, ops->cn_deg[i])];
        p->cn_nodes[i] = c.d_nodes;
        p->cn_count[i] = c.count;
        if (int rc = starts_of(c, h->h_sc, &p->cn_starts[i])) return rc;
    }
    for (size_t i = 0; i < ops->vn_deg.size(); ++i) {
        const NodeClass& c = h->vn_classes[class_index(h->vn_classes, ops->vn_deg[i])];
        p->vn_nodes[i] = c.d_nodes;
        p->vn_count[i] = c.count;
        if (int rc = starts_of(c, h->h_sv, &p->vn_starts[i])) return rc;
    }
    return IBLDPC_OK;
}

// Called at the end of ibldpc_set_luts (packed-nibble family, default tail-pair thresholds): picks the instantiated
// degree set of this code, expands every phase image on the host and uploads them.  Leaves h->phase == nullptr
// (per-class launches) when the code's degree sets are not instantiated.
int phase_prepare(ibldpc_decoder* h)
{
    phase_free(h);
    if (!h->nib || !h->use_pair || !h->use_phase) return IBLDPC_OK;
    // Measured on B200 (profiles/README.md, round 2): the fused kernels sit on the same look-up-pipe / issue ceiling as
    // the per-class ones, so they win where a phase is several launches with short classes (802.11n: 3.01 -> 3.51
    // Gbit/s) and lose a few per cent where one class dominates a phase (DVB-S2 rate 1/2: 4.15 -> 3.95, regular (3,6):
    // 6.01 -> 5.82).  Default: on for the 802.11n degree sets; IBLDPC_PHASE=1 turns them on for every instantiated set.
    struct Cand { const PhaseSetOps* ops; bool by_default; };
    const Cand candidates[] = {{phase_ops_wlan(), true}, {phase_ops_dvbs2(), false}, {phase_ops_reg36(), false}};
    const bool force = getenv("IBLDPC_PHASE") != nullptr && atoi(getenv("IBLDPC_PHASE")) != 0;
    const PhaseSetOps* ops = nullptr;
    h->phase_default = false;
    for (const Cand& c : candidates)
        if (same_degrees(c.ops->cn_deg, h->cn_classes) && same_degrees(c.ops->vn_deg, h->vn_classes)) {
            ops = c.ops;                                   // images are built for every instantiated set (per-frame early
            h->phase_default = c.by_default || force;      // termination runs on them); the plain decode uses them by default
        }                                                  // only where they win
    if (!ops) return IBLDPC_OK;
    if (int rc = build_phase_images(h, ops, &h->phase)) return rc;
    // the (3,6) set a second time through the three-input tables: per-frame early termination runs on these images
    // (check-node pass 547 -> 430 us, variable-node pass 450 -> 350 us at 65536 frames); the plain images stay the ones of
    // the small-batch and mid-range paths, where a 192 KB image per launch would cost more than the look-ups it saves
    if (ops == phase_ops_reg36() && !h->h_cn3.empty() && !h->h_vn3.empty() && getenv("IBLDPC_NO_PF_TRIPLE") == nullptr)
        if (int rc = build_phase_images(h, phase_ops_reg36_tri(), &h->phase_tri)) return rc;
    return IBLDPC_OK;
}

bool phase_available(const ibldpc_decoder* h) { return h->phase != nullptr; }

const PhaseSetOps* phase_ops_of(const ibldpc_decoder* h) { return pf_images(h) ? pf_images(h)->ops : nullptr; }

// image pointer, node lists and dynamic shared memory of one phase: mode kPhaseCn with table block `index`,
// kPhaseVn with iteration `index`, kPhaseOut with the decision tables of iteration `index` (per-frame mode)
void phase_fill_args(const ibldpc_decoder* h, int mode, int index, PhaseArgs& q, size_t* smem)
{
    const PhaseImages* p = pf_images(h);
    const PhaseSetOps* ops = p->ops;
    if (mode == kPhaseCn) {
        q.image = p->d_images + (size_t)index * p->cn_bytes;
        *smem = p->cn_bytes;
        for (int i = 0; i < ops->cn_layout.n; ++i) {
            q.nodes[i] = p->cn_nodes[i]; q.starts[i] = p->cn_starts[i]; q.n_nodes[i] = p->cn_count[i];
        }
        return;
    }
    if (mode == kPhaseVn) {
        q.image = p->d_images + (size_t)p->imax * p->cn_bytes + (size_t)index * p->vn_bytes;
        *smem = p->vn_bytes;
    } else {
        q.image = p->d_images + (size_t)p->imax * (p->cn_bytes + p->vn_bytes);
        q.image_stride = (long long)p->out_bytes;       // the kernel adds a.it * stride
        *smem = p->out_bytes;
    }
    for (int i = 0; i < ops->vn_layout.n; ++i) {
        q.nodes[i] = p->vn_nodes[i]; q.starts[i] = p->vn_starts[i]; q.n_nodes[i] = p->vn_count[i];
    }
}

static int set_attributes(PhaseImages* p)
{
    const PhaseSetOps* ops = p->ops;
    if (p->occ_checked) return IBLDPC_OK;
    const struct { const void* k; size_t smem; } ks[] = {{(const void*)ops->cn_kernel[0], p->cn_bytes}, {(const void*)ops->cn_kernel[1], p->cn_bytes},
                                                         {(const void*)ops->vn_kernel, p->vn_bytes}, {(const void*)ops->out_kernel, p->out_bytes},
                                                         {(const void*)ops->cn_pf_kernel[0], p->cn_bytes}, {(const void*)ops->cn_pf_kernel[1], p->cn_bytes},
                                                         {(const void*)ops->vn_pf_kernel, p->vn_bytes},
                                                         {(const void*)ops->pf_decide_kernel, p->out_bytes}};
    for (auto& e : ks) {
        // the per-frame check-node kernel keeps its syndrome accumulator behind the image: allow the full 227 KB
        const int limit = e.k == (const void*)ops->cn_pf_kernel[0] ? 227 * 1024 - 1024 : (int)e.smem;
        IBLDPC_CK(cudaFuncSetAttribute(e.k, cudaFuncAttributeMaxDynamicSharedMemorySize, limit));
        int occ = 0;
        IBLDPC_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, e.k, kPhaseThreads, e.smem));
        if (occ < 1) return fail_msg(IBLDPC_E_CUDA, "fused per-phase kernel does not fit on an SM");
    }
    p->occ_checked = 1;
    return IBLDPC_OK;
}

int phase_set_attributes(ibldpc_decoder* h)
{
    if (int rc = set_attributes(h->phase)) return rc;
    return h->phase_tri ? set_attributes(h->phase_tri) : IBLDPC_OK;
}

namespace {

struct PhaseProf {
    ibldpc_decoder* h;
    cudaStream_t st;
    int idx = -1;
    int begin(int phase)
    {
        if (!h->profiling) return IBLDPC_OK;
        PhaseEvent ev;
        ev.phase = phase;
        IBLDPC_CK(cudaEventCreate(&ev.a));
        IBLDPC_CK(cudaEventCreate(&ev.b));
        IBLDPC_CK(cudaEventRecord(ev.a, st));
        h->events.push_back(ev);
        idx = (int)h->events.size() - 1;
        return IBLDPC_OK;
    }
    int end()
    {
        if (!h->profiling || idx < 0) return IBLDPC_OK;
        IBLDPC_CK(cudaEventRecord(h->events[idx].b, st));
        return IBLDPC_OK;
    }
};

}  // namespace

// One decode with the fused kernels.  `a` carries graph pointers, buffers (packed channel values in a.ch, message
// array, output), pitches, flags and the iteration control of decode_ib_n4; the launches go to `st`.
int decode_ib_phase(ibldpc_decoder* h, const IbArgs& a, int imax, int early, cudaStream_t st)
{
    PhaseImages* p = h->phase;
    const PhaseSetOps* ops = p->ops;
    if (int rc0 = phase_set_attributes(h)) return rc0;
    auto grid_for = [&](const PhaseLayoutRt& L, const int* counts) {
        long long chunks = 0;   // groups of 32 items: below one group per CTA there is nothing to share
        for (int i = 0; i < L.n; ++i) {
            const long long tiles = ((long long)a.pitch + 128LL * L.cls[i].vec - 1) / (128LL * L.cls[i].vec);
            chunks += ((long long)counts[i] * tiles + 31) / 32;
        }
        return (int)std::max<long long>(1, std::min<long long>(h->sm_count, chunks));
    };
    PhaseArgs base{};
    base.a = a;
    base.a.early = early;
    base.a.imax = imax;
    PhaseProf prof{h, st};
    int rc;
    auto launch_cn = [&](int it) -> int {
        PhaseArgs q = base;
        q.a.it = it;
        q.a.iter0 = it < 0;
        q.image = p->d_images + (size_t)(it + 1) * p->cn_bytes;
        for (int i = 0; i < ops->cn_layout.n; ++i) {
            q.nodes[i] = p->cn_nodes[i]; q.starts[i] = p->cn_starts[i]; q.n_nodes[i] = p->cn_count[i];
        }
        if ((rc = prof.begin(it < 0 ? 2 : 0))) return rc;
        const int grid = grid_for(ops->cn_layout, p->cn_count);
        ops->cn_kernel[early ? 1 : 0]<<<grid, kPhaseThreads, p->cn_bytes, st>>>(q);
        h->last_launches++; h->last_grid = grid; h->last_smem = (int)p->cn_bytes;
        return prof.end();
    };
    auto launch_vn = [&](int it, bool decide) -> int {
        PhaseArgs q = base;
        q.a.it = it;
        q.a.iter0 = 0;
        if (decide) {
            q.image = p->d_images + (size_t)p->imax * (p->cn_bytes + p->vn_bytes);
            q.image_stride = (long long)p->out_bytes;
        } else {
            q.image = p->d_images + (size_t)p->imax * p->cn_bytes + (size_t)it * p->vn_bytes;
        }
        for (int i = 0; i < ops->vn_layout.n; ++i) {
            q.nodes[i] = p->vn_nodes[i]; q.starts[i] = p->vn_starts[i]; q.n_nodes[i] = p->vn_count[i];
        }
        if ((rc = prof.begin(decide ? 2 : 1))) return rc;
        const PhaseLayoutRt& L = decide ? ops->out_layout : ops->vn_layout;
        const int grid = grid_for(L, p->vn_count);
        const size_t smem = decide ? p->out_bytes : p->vn_bytes;
        (decide ? ops->out_kernel : ops->vn_kernel)<<<grid, kPhaseThreads, smem, st>>>(q);
        h->last_launches++;
        return prof.end();
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

// One decode of a small batch (B <= kLaneModeMaxFrames) in ONE cooperative launch (ib_coop_phase_kernel): lane = (node, word),
// the image of every phase brought in by TMA while the grid barrier before the phase completes.
int decode_ib_coop_phase(ibldpc_decoder* h, const IbArgs& a, long long B, int imax, int early, cudaStream_t st)
{
    PhaseImages* p = h->phase;
    const PhaseSetOps* ops = p->ops;
    const size_t smem = std::max(p->cn_bytes, std::max(p->vn_bytes, p->out_bytes));
    const int NT = ops->coop_threads;
    CoopPhaseKernel k = ops->coop_kernel[early ? 1 : 0];
    if (!p->coop_checked) {
        for (CoopPhaseKernel kk : {ops->coop_kernel[0], ops->coop_kernel[1]}) {
            IBLDPC_CK(cudaFuncSetAttribute((const void*)kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ = 0;
            IBLDPC_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)kk, NT, smem));
            if (occ < 1) return fail_msg(IBLDPC_E_CUDA, "cooperative per-phase kernel does not fit on an SM");
        }
        p->coop_checked = 1;
    }
    CoopPhaseArgs q{};
    q.a = a;
    q.a.early = early;
    q.a.imax = imax;
    q.cn_images = p->d_images;
    q.vn_images = p->d_images + (size_t)p->imax * p->cn_bytes;
    q.out_images = p->d_images + (size_t)p->imax * (p->cn_bytes + p->vn_bytes);
    // warps the busiest class can use: lane mode (B <= kLaneModeMaxFrames) 32 / ceil(B/8) nodes per warp step, above one
    // (node, tile) item per warp step
    const bool lanes = B <= kLaneModeMaxFrames;
    const long long wpn = (B + 7) / 8, per_warp = lanes ? 32 / wpn : 1;
    auto warps_of = [&](int count, int vec) -> long long {
        if (lanes) return (count + per_warp - 1) / per_warp;
        return (long long)count * (((long long)a.pitch + 128LL * vec - 1) / (128LL * vec));
    };
    long long warps = 1;
    for (int i = 0; i < ops->cn_layout.n; ++i) {
        q.cn_nodes[i] = p->cn_nodes[i]; q.cn_starts[i] = p->cn_starts[i]; q.cn_count[i] = p->cn_count[i];
        warps = std::max(warps, warps_of(p->cn_count[i], ops->cn_layout.cls[i].vec));
    }
    for (int i = 0; i < ops->vn_layout.n; ++i) {
        q.vn_nodes[i] = p->vn_nodes[i]; q.vn_starts[i] = p->vn_starts[i]; q.vn_count[i] = p->vn_count[i];
        warps = std::max(warps, warps_of(p->vn_count[i], ops->vn_layout.cls[i].vec));
    }
    const int grid = (int)std::max<long long>(1, std::min<long long>(h->sm_count, (warps + NT / 32 - 1) / (NT / 32)));
    PhaseProf prof{h, st};
    if (int rc = prof.begin(2)) return rc;
    void* params[] = {&q};
    IBLDPC_CK(cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(NT), params, smem, st));
    h->last_launches++; h->last_grid = grid; h->last_smem = (int)smem;
    if (int rc = prof.end()) return rc;
    return IBLDPC_OK;
}

}  // namespace ibldpc
