// ibldpc.cu -- host side of libibldpc.so: C ABI (include/ibldpc.h), table upload, launch
// sequencing of the flooding schedule, pinned-host pipeline.  Kernels live in ib_kernels.cuh
// and llr_kernels.cuh.  Compiled for sm_100a only; there is no CPU fallback anywhere.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "ibldpc_internal.h"
#include "ib_kernels.cuh"
#include "ib_kernels_n4.cuh"
#include "ib_coop_n4.cuh"
#include "ib_triple_n4.cuh"
#include "llr_kernels.cuh"
#include "kernel_tables.h"

using namespace ibldpc;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(IBLDPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
    } while (0)

}  // namespace

namespace ibldpc {
// error sink shared with encoder.cu
int fail_msg(int code, const std::string& msg) { return fail(code, msg); }
}  // namespace ibldpc

namespace {

template <typename T>
int upload(T** dst, const T* src, size_t n)
{
    CK(cudaMalloc((void**)dst, std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return IBLDPC_OK;
}

int build_classes(const std::vector<int>& deg, std::vector<NodeClass>& out)
{
    std::map<int, std::vector<int>> by;
    for (int i = 0; i < (int)deg.size(); ++i) by[deg[i]].push_back(i);
    for (auto& kv : by) {
        NodeClass c;
        c.degree = kv.first;
        c.count = (int)kv.second.size();
        int rc = upload(&c.d_nodes, kv.second.data(), kv.second.size());
        if (rc) return rc;
        out.push_back(c);
    }
    return IBLDPC_OK;
}

int ensure(void** p, size_t* have, size_t need, bool zero = false)
{
    if (*have >= need && *p) return IBLDPC_OK;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e != cudaSuccess) return fail(IBLDPC_E_NOMEM, std::string("cudaMalloc of ") + std::to_string(need) + " bytes: " + cudaGetErrorString(e));
    if (zero) CK(cudaMemset(*p, 0, need));
    *have = need;
    return IBLDPC_OK;
}

int ensure_ws_common(ibldpc_decoder* h, Workspace& w)
{
    if (!w.flags) {
        CK(cudaMalloc((void**)&w.flags, sizeof(int) * (kMaxIter + 1)));   // [0] = input-range flag, [1..] = per-pass syndrome flags
        CK(cudaMemset(w.flags, 0, sizeof(int) * (kMaxIter + 1)));
        CK(cudaMalloc((void**)&w.inum, sizeof(int)));
    }
    (void)h;
    return IBLDPC_OK;
}

int occupancy_of(ibldpc_decoder* h, const void* fn, int smem, int* occ_out, int threads = kThreads)
{
    auto key = std::make_pair(fn, smem);
    auto it = h->occ_cache.find(key);
    if (it == h->occ_cache.end()) {
        int occ;
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads, smem));
        if (occ < 1) return fail(IBLDPC_E_CUDA, "kernel does not fit on an SM");
        h->occ_cache[key] = occ;
        *occ_out = occ;
    } else {
        *occ_out = it->second;
    }
    return IBLDPC_OK;
}

// IB fast path: grid.x CTAs per tile group so that grid.x * tile_groups fills the resident slots
int grid_for(ibldpc_decoder* h, const void* fn, int smem, int tile_groups, int nodes_per_step, int n_nodes, int* out,
             double share = 1.0, int threads = kThreads)
{
    int occ;
    int rc = occupancy_of(h, fn, smem, &occ, threads);
    if (rc) return rc;
    // floor: one CTA more than the resident slots would run as a second wave
    const long long cap = std::max<long long>(1, (long long)(share * ((long long)occ * h->sm_count / tile_groups)));
    const long long want = ((long long)n_nodes + nodes_per_step - 1) / nodes_per_step;
    *out = (int)std::max<long long>(1, std::min(want, cap));
    return IBLDPC_OK;
}

// persistent grid: resident CTAs per SM x SM count, capped by the work available
int grid_for(ibldpc_decoder* h, const void* fn, int smem, long long items, int* out)
{
    auto key = std::make_pair(fn, smem);
    auto it = h->occ_cache.find(key);
    int occ;
    if (it == h->occ_cache.end()) {
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kThreads, smem));
        if (occ < 1) return fail(IBLDPC_E_CUDA, "kernel does not fit on an SM");
        h->occ_cache[key] = occ;
    } else {
        occ = it->second;
    }
    long long want = (items + kWarpsPerCta - 1) / kWarpsPerCta;
    long long cap = (long long)occ * h->sm_count;
    *out = (int)std::max<long long>(1, std::min(want, cap));
    return IBLDPC_OK;
}

// ------------------------------------------------------------------------------------------
// Host-side planning (pure functions, also exported for the CPU test-suite: ibldpc_plan_geometry,
// ibldpc_host_chunk_schedule)
// ------------------------------------------------------------------------------------------
// Launch geometry of one degree class of the packed-nibble kernels: the grid is (CTAs per tile group) x (tile
// groups), a CTA's warps cover 2^tpc_log2 consecutive tiles of (warps >> tpc_log2) nodes.
// floor(resident slots / tile groups) CTAs per group can leave many slots empty (802.11n, B=100096: 25 groups on
// 148 slots -> 125 CTAs), so the number of tiles per CTA is chosen to minimise the makespan in node steps:
// waves x ceil(node steps / CTAs per group) / (fraction of warps that own a real tile); ties go to the wider CTA
// footprint (longer contiguous runs per row).  `widest_only` = the pre-planner rule (A/B switch IBLDPC_NO_PLAN).
void plan_geometry(long long slots, int warps, int tiles, int n_nodes, bool widest_only, int* tpc_log2, int* tile_groups,
                   int* grid_x)
{
    double best = 0;
    int best_tpc = -1;
    for (int tpc = 5; tpc >= 0; --tpc) {      // up to all 32 warps of a 1024-thread CTA on consecutive tiles of one node
        if ((warps >> tpc) < 1) continue;
        if (warps & ((1 << tpc) - 1)) continue;      // 24-warp CTAs (768 threads): the warps must split into whole groups of 2^tpc tiles
        if (tpc > 0 && (1 << (tpc - 1)) >= tiles) continue;      // wider than the row: only wasted warps
        if (widest_only && best_tpc >= 0) break;
        const long long tg = (tiles + (1 << tpc) - 1) >> tpc;
        const long long nsteps = (n_nodes + (warps >> tpc) - 1) / (warps >> tpc);
        const long long per = std::max<long long>(1, std::min(nsteps, slots / tg));
        const long long waves = (per * tg + slots - 1) / slots;
        const double used = (double)tiles / (double)(tg << tpc);
        const double cost = (double)(waves * ((nsteps + per - 1) / per)) / used;
        if (best_tpc < 0 || cost < best * 0.999) { best = cost; best_tpc = tpc; }
    }
    *tpc_log2 = best_tpc;
    *tile_groups = (tiles + (1 << best_tpc) - 1) >> best_tpc;
    const long long nsteps = (n_nodes + (warps >> best_tpc) - 1) / (warps >> best_tpc);
    *grid_x = (int)std::max<long long>(1, std::min(nsteps, std::max<long long>(1, slots / *tile_groups)));
}

// Chunk schedule of the two-slot copy/decode pipeline of ibldpc_decode_ib_host.  Auto (host_chunk == 0): equal
// chunks of at most ~256 MiB of channel values (at least two when B >= 8192), with the first and the last chunk
// split 1/4 + 3/4 and 3/4 + 1/4 so that only a quarter chunk of copy-in and of copy-out is not overlapped with
// decoding.  Measured on B200, C1, B=65536: 8 x 8192 frames 4.2, 2 x 32768 4.64, ramped 5.03 Gbit/s end to end
// (the kernels lose efficiency on small batches, large chunks expose their first copy-in / last copy-out).
// An explicit chunk size gives equal chunks; early termination is a property of the whole call: one chunk.
std::vector<int64_t> host_chunk_schedule(int64_t B, int64_t n_var, int64_t host_chunk, bool early_term)
{
    std::vector<int64_t> widths;
    if (early_term) {
        widths.push_back(B);
    } else if (host_chunk > 0) {
        const int64_t c = std::max<int64_t>(16, host_chunk / 16 * 16);
        for (int64_t off = 0; off < B; off += c) widths.push_back(std::min<int64_t>(c, B - off));
    } else {
        const int64_t target = std::max<int64_t>(512, ((256LL << 20) / std::max<int64_t>(n_var, 1)) / 512 * 512);
        int64_t n_chunks = (B + target - 1) / target;
        if (n_chunks < 2 && B >= 8192) n_chunks = 2;
        if (n_chunks < 2) {
            widths.push_back(B);
        } else {
            const int64_t c = ((B + n_chunks - 1) / n_chunks + 511) / 512 * 512;
            const int64_t q = std::max<int64_t>(512, (c / 4) / 512 * 512);
            int64_t left = B;
            auto take = [&](int64_t wdt) { wdt = std::min(wdt, left); if (wdt > 0) { widths.push_back(wdt); left -= wdt; } };
            take(q);
            take(c - q);
            while (left > c) take(c);
            if (left > q) take(std::max<int64_t>(512, (left - q) / 512 * 512));   // every chunk but the last: whole 512-frame tiles
            take(left);
        }
    }
    return widths;
}

struct Prof {
    ibldpc_decoder* h;
    cudaStream_t st;
    int idx = -1;
    int begin(int phase)
    {
        if (!h->profiling) return IBLDPC_OK;
        PhaseEvent ev;
        ev.phase = phase;
        CK(cudaEventCreate(&ev.a));
        CK(cudaEventCreate(&ev.b));
        CK(cudaEventRecord(ev.a, st));
        h->events.push_back(ev);
        idx = (int)h->events.size() - 1;
        return IBLDPC_OK;
    }
    int end()
    {
        if (!h->profiling || idx < 0) return IBLDPC_OK;
        CK(cudaEventRecord(h->events[idx].b, st));
        return IBLDPC_OK;
    }
};

void clear_events(ibldpc_decoder* h)
{
    for (auto& e : h->events) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    h->events.clear();
}

int decode_ib_n4(ibldpc_decoder* h, Workspace& w, const uint8_t* ch, long long pitch8, long long B, int imax, int early,
                 uint8_t* out, cudaStream_t st, bool ch_packed);

// ------------------------------------------------------------------------------------------
// IB decode on padded device buffers (pitch multiple of 16, pointers 16-byte aligned)
// ------------------------------------------------------------------------------------------
int decode_ib_padded(ibldpc_decoder* h, Workspace& w, const uint8_t* ch, long long pitch, long long B, int imax,
                     int early, uint8_t* out, cudaStream_t st, bool ch_packed = false)
{
    h->last_stream = st;
    h->last_ws = (int)(&w - &h->ws[0]);
    if (h->fast && h->nib) return decode_ib_n4(h, w, ch, pitch, B, imax, early, out, st, ch_packed);
    int rc = ensure_ws_common(h, w);
    if (rc) return rc;
    if (ch_packed) return fail(IBLDPC_E_STATE, "packed channel buffers need the packed-nibble kernel family");
    size_t need = (size_t)h->E * (size_t)pitch;
    {
        void* p = w.msg;
        rc = ensure(&p, &w.msg_bytes, need);
        w.msg = (uint8_t*)p;
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(w.flags + 1, 0, sizeof(int) * (size_t)std::max(imax, 1), st));   // flags[0] (input range) is sticky until read
    {   // These families use the channel values as table indices as they are: work on a sanitised copy (values
        // >= |T_channel| clamped and reported through flags[0]) so that no look-up can leave its table.
        void* p = w.ch4;
        rc = ensure(&p, &w.ch4_bytes, (size_t)h->N * (size_t)pitch);
        w.ch4 = (uint8_t*)p;
        if (rc) return rc;
        const long long n = (long long)h->N * (pitch / 16);
        const int grid = (int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 16);
        clamp_u8_kernel<<<grid, 256, 0, st>>>(ch, w.ch4, n, h->Tc, w.flags);
        h->last_launches = 0;
        ch = w.ch4;
    }
    IbArgs a{};
    a.sc = h->d_sc; a.deg_c = h->d_dc; a.sv = h->d_sv; a.deg_v = h->d_dv; a.tv = h->d_tv; a.vidx = h->d_vidx;
    a.n_var = h->N; a.n_chk = h->M;
    a.ch = ch; a.msg = w.msg; a.out = out;
    a.pitch = (uint32_t)pitch; a.B = (int)B; a.tiles = (int)((pitch + 511) / 512);
    a.tpc_log2 = a.tiles >= 8 ? 3 : a.tiles > 2 ? 2 : a.tiles == 2 ? 1 : 0;
    const int tile_groups = (a.tiles + (1 << a.tpc_log2) - 1) >> a.tpc_log2;
    const int nps = kWarpsPerCta >> a.tpc_log2;
    a.T = h->T; a.Tc = h->Tc; a.tshift = h->tshift;
    a.flags = w.flags + 1; a.inum = w.inum; a.early = early; a.imax = imax;
    a.DC = h->DC; a.DV = h->DV; a.xp_col = -1;
    const int T = h->T, Tc = h->Tc, TT = T * T;
    h->last_launches = 0;
    if (h->t32) return decode_ib_t32(h, a, imax, early, st);
    Prof prof{h, st};

    // Irregular codes: the degree classes of one phase are independent, so their kernels CAN run
    // concurrently -- class 0 on the caller's stream, the others on auxiliary streams forked from and
    // joined back to it by events -- each with a share of the resident CTA slots proportional to its
    // estimated work.  Measured on B200 (round 1) this is SLOWER than launching the classes one after
    // the other (WLAN n=1296: 1.52 vs 1.78 Gbit/s, DVB-S2-like: 1.72 vs 2.05): the static split leaves the
    // phase waiting for its slowest class.  Kept behind IBLDPC_CONCURRENT_CLASSES=1 for further tuning.
    const bool concurrent = std::max(h->cn_classes.size(), h->vn_classes.size()) > 1 && getenv("IBLDPC_CONCURRENT_CLASSES");
    if (concurrent && !w.fork_ev) {
        CK(cudaEventCreateWithFlags(&w.fork_ev, cudaEventDisableTiming));
        for (int i = 0; i < 7; ++i) {
            CK(cudaStreamCreateWithFlags(&w.aux[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&w.join_ev[i], cudaEventDisableTiming));
        }
    }
    auto cn_cost = [](int d) { return 2.0 * (d - 2) + 0.5 * (d - 1) * (d - 2) + 0.5 * d + 1.0; };
    auto vn_cost = [](int d) { return (d - 1) + 0.5 * d * (d - 1) + 0.25 * (2 * d + 1) + 1.0; };
    auto class_stream = [&](size_t i) -> cudaStream_t { return (concurrent && i > 0) ? w.aux[(i - 1) % 7] : st; };
    auto fork_phase = [&](size_t nclasses) -> int {
        if (concurrent && nclasses > 1) {
            CK(cudaEventRecord(w.fork_ev, st));
            for (size_t i = 1; i < nclasses; ++i) CK(cudaStreamWaitEvent(class_stream(i), w.fork_ev, 0));
        }
        return IBLDPC_OK;
    };
    auto join_class = [&](size_t i) -> int {
        if (concurrent && i > 0) {
            CK(cudaEventRecord(w.join_ev[(i - 1) % 7], class_stream(i)));
            CK(cudaStreamWaitEvent(st, w.join_ev[(i - 1) % 7], 0));
        }
        return IBLDPC_OK;
    };
    if (h->fast) {
        // Each degree class only stages the tables its chains reach (stages 0..d-3 for a check of
        // degree d, 0..d-2 / 0..d-1 for a variable node), so low-degree classes of an irregular code
        // get small shared-memory footprints and high occupancy.
        auto words = [](int cols) { return std::max(1, (cols + 3) / 4); };
        auto launch_cn = [&](int it) -> int {
            IbArgs b = a;
            b.it = it; b.iter0 = (it < 0);
            const int blk = it + 1;   // table block: 0 = iteration-0 tables
            b.lut = h->d_cn8 + (size_t)blk * (h->DC - 2) * TT;
            b.match = h->match ? h->d_mc8 + (size_t)blk * h->DC * T : nullptr;
            b.dmax_match = h->DC;
            int r = prof.begin(it < 0 ? 2 : 0);
            if (r) return r;
            if ((r = fork_phase(h->cn_classes.size()))) return r;
            double total_work = 0;
            for (auto& c : h->cn_classes) total_work += c.count * cn_cost(c.degree);
            size_t ci_run = 0;
            for (auto& c : h->cn_classes) {
                cudaStream_t cs = class_stream(ci_run);
                const double share = concurrent ? c.count * cn_cost(c.degree) / total_work : 1.0;
                b.nst = c.degree - 2;
                const bool explicit_match = h->match && b.nst == 0;   // matching is folded into the last stage otherwise
                b.W = words(b.nst + (explicit_match ? 1 : 0));
                b.nrows = std::max(TT, explicit_match ? c.degree * T : 0);
                if (h->match) b.dmax_match = c.degree;
                // Measured on B200: the tail-pair variant needs 64 KB of shared memory and 80 registers (3 CTAs/SM);
                // it wins for d_c >= 7 (WLAN, DVB-S2: -3 % CN time) and loses for d_c = 6 (+5 %), so it is used
                // from degree 7 on (IBLDPC_PAIR_MIN_DEGREE overrides, IBLDPC_NO_PAIR disables).
                const bool pair = h->use_pair && c.degree >= h->pair_min_degree && h->d_cn_pair != nullptr;
                b.xp_col = -1;
                int smem_main = b.nrows * b.W * 128;
                smem_main += stage_scratch_bytes(b.nst, T, h->match ? b.dmax_match : 0);   // [tables][scratch][pair rows]
                if (pair) {
                    const size_t ci = (size_t)(&c - &h->cn_classes[0]);
                    b.pair = h->d_cn_pair + ((size_t)blk * h->cn_classes.size() + ci) * (size_t)TT * 8;
                    b.pair_off = (uint32_t)smem_main;
                    smem_main += TT * 8 * kPairSlots;
                    b.xp_col = c.degree - 5;    // -1 for degree 4: no stage feeds the pair row
                }
                NodeKernel k = cn_fast_kernel_for(c.degree, explicit_match, early != 0, pair);
                const int smem = smem_main;
                int grid;
                r = grid_for(h, (const void*)k, smem, tile_groups, nps, c.count, &grid, share);
                if (r) return r;
                k<<<dim3(grid, tile_groups), kThreads, smem, cs>>>(b, c.d_nodes, c.count);
                h->last_launches++; h->last_grid = grid * tile_groups; h->last_smem = smem;
                if ((r = join_class(ci_run++))) return r;
            }
            return prof.end();
        };
        auto launch_vn = [&](int it, bool decide) -> int {
            IbArgs b = a;
            b.it = it; b.iter0 = 0;
            if (decide) {
                b.lut = h->d_vn8; b.vn_it_stride = (long long)h->DV * TT; b.match = nullptr;
            } else {
                b.lut = h->d_vn8 + (size_t)it * h->DV * TT;
                b.match = h->match ? h->d_mv8 + (size_t)it * h->DV * T : nullptr;
            }
            int r = prof.begin(decide ? 2 : 1);
            if (r) return r;
            if ((r = fork_phase(h->vn_classes.size()))) return r;
            double total_work = 0;
            for (auto& c : h->vn_classes) total_work += c.count * vn_cost(c.degree);
            size_t ci_run = 0;
            for (auto& c : h->vn_classes) {
                cudaStream_t cs = class_stream(ci_run);
                const double share = concurrent ? c.count * vn_cost(c.degree) / total_work : 1.0;
                const bool m = b.match != nullptr;
                b.nst = decide ? c.degree : c.degree - 1;
                b.W = words(b.nst);                   // matching is folded into the last stage (stage_tables)
                b.nrows = TT;
                b.dmax_match = m ? c.degree : 0;
                const int smem = b.nrows * b.W * 128 + stage_scratch_bytes(b.nst, T, b.dmax_match);
                NodeKernel k = vn_fast_kernel_for(c.degree, decide, false);
                int grid;
                r = grid_for(h, (const void*)k, smem, tile_groups, nps, c.count, &grid, share);
                if (r) return r;
                k<<<dim3(grid, tile_groups), kThreads, smem, cs>>>(b, c.d_nodes, c.count);
                h->last_launches++;
                if ((r = join_class(ci_run++))) return r;
            }
            return prof.end();
        };
        if ((rc = launch_cn(-1))) return rc;
        for (int it = 0; it < imax - 1; ++it) {
            if ((rc = launch_vn(it, false))) return rc;
            if ((rc = launch_cn(it))) return rc;
        }
        if ((rc = launch_vn(0, true))) return rc;
    } else {
        a.lut_all = h->d_cn8;
        const dim3 blk(128);
        auto grid2 = [&](int nodes) { return dim3((unsigned)((pitch + 127) / 128), (unsigned)std::min(nodes, 4096)); };
        auto launch_cn = [&](int it) -> int {
            IbArgs b = a;
            b.it = it < 0 ? 0 : it; b.iter0 = (it < 0);
            b.lut_all = h->d_cn8; b.match_all = h->match ? h->d_mc8 : nullptr;
            int r = prof.begin(it < 0 ? 2 : 0);
            if (r) return r;
            ib_cn_generic_kernel<<<grid2(h->M), blk, 0, st>>>(b);
            h->last_launches++;
            return prof.end();
        };
        auto launch_vn = [&](int it, bool decide) -> int {
            IbArgs b = a;
            b.it = it; b.lut_all = h->d_vn8; b.match_all = h->match ? h->d_mv8 : nullptr;
            int r = prof.begin(decide ? 2 : 1);
            if (r) return r;
            if (decide) ib_vn_generic_kernel<true><<<grid2(h->N), blk, 0, st>>>(b);
            else ib_vn_generic_kernel<false><<<grid2(h->N), blk, 0, st>>>(b);
            h->last_launches++;
            return prof.end();
        };
        (void)Tc;
        if ((rc = launch_cn(-1))) return rc;
        for (int it = 0; it < imax - 1; ++it) {
            if ((rc = launch_vn(it, false))) return rc;
            if ((rc = launch_cn(it))) return rc;
        }
        if ((rc = launch_vn(0, true))) return rc;
    }
    CK(cudaGetLastError());
    return IBLDPC_OK;
}

// ------------------------------------------------------------------------------------------
// IB decode, packed-nibble fast path: `ch` / `out` are the uint8 buffers of decode_ib_padded
// (pitch multiple of 16); messages and channel values travel as nibbles inside.
// ------------------------------------------------------------------------------------------
int decode_ib_n4(ibldpc_decoder* h, Workspace& w, const uint8_t* ch, long long pitch8, long long B, int imax, int early,
                 uint8_t* out, cudaStream_t st, bool ch_packed)
{
    int rc = ensure_ws_common(h, w);
    if (rc) return rc;
    const long long pitch4 = ((B + 1) / 2 + 15) / 16 * 16;
    {
        void* p = w.msg;
        rc = ensure(&p, &w.msg_bytes, (size_t)h->E * (size_t)pitch4);
        w.msg = (uint8_t*)p;
        if (rc) return rc;
        p = w.ch4;
        rc = ensure(&p, &w.ch4_bytes, (size_t)h->N * (size_t)pitch4);
        w.ch4 = (uint8_t*)p;
        if (rc) return rc;
    }
    CK(cudaMemsetAsync(w.flags + 1, 0, sizeof(int) * (size_t)std::max(imax, 1), st));   // flags[0] (input range) is sticky until read
    h->last_launches = 0;
    Prof prof{h, st};
    if (!ch_packed) {   // ch_packed: the caller already filled w.ch4 (packed host path)
        const long long nwords = (long long)h->N * (pitch4 / 4);
        const int grid = (int)std::min<long long>((nwords + 255) / 256, (long long)h->sm_count * 16);
        if ((rc = prof.begin(2))) return rc;
        pack_n4_kernel<true><<<grid, 256, 0, st>>>(ch, w.ch4, h->N, B, pitch8, (uint32_t)pitch4, h->T, w.flags);
        h->last_launches++;
        if ((rc = prof.end())) return rc;
    }
    IbArgs a{};
    a.sc = h->d_sc; a.deg_c = h->d_dc; a.sv = h->d_sv; a.deg_v = h->d_dv; a.tv = h->d_tv; a.vidx = h->d_vidx;
    a.n_var = h->N; a.n_chk = h->M;
    a.ch = w.ch4; a.msg = w.msg; a.out = out;
    a.pitch = (uint32_t)pitch4; a.out_pitch = (uint32_t)pitch8; a.B = (int)B;
    a.T = h->T; a.Tc = h->Tc; a.tshift = h->tshift;
    a.flags = w.flags + 1; a.inum = w.inum; a.early = early; a.imax = imax;
    a.DC = h->DC; a.DV = h->DV; a.xp_col = -1;
    const int T = h->T, TT = T * T;
    // ---- opt-in: per-frame early termination with frame compaction (ib_perframe.cu)
    if (h->pf_request) return decode_ib_perframe(h, w, a, B, imax, h->pf_inum, st);
    // ---- small batches of the instantiated degree sets: one cooperative launch over the TMA-staged phase images
    if (h->phase && !h->no_coop_phase && B <= h->coop_max_frames &&
        (B <= kLaneModeMaxFrames || h->coop_max_from_env || (long long)h->E * pitch4 <= (32LL << 20))) {
        if (h->coop_supported < 0) {
            int v = 0;
            CK(cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->device));
            h->coop_supported = v;
        }
        if (h->coop_supported) return decode_ib_coop_phase(h, a, B, imax, early, st);
    }
    // ---- small batches of regular codes: the whole decode in one cooperative launch (ib_coop_n4.cuh)
    // (worth it while a phase is short: measured break-even near 100 MB of packed messages -- DVB-S2 n=64800 wins at
    // B=512 (58 MB: 9.3 -> 6.9 ms) and loses at B=2048 (232 MB); C1 and 802.11n win up to the 4096-frame limit)
    // (codes without phase images, or IBLDPC_NO_COOP_PHASE=1: the instantiated sets took the branch above or go on to
    // the fused per-phase kernels)
    const bool coop_fits = (h->phase == nullptr || h->no_coop_phase) && B <= h->coop_max_frames && pitch4 <= 8 * 256 &&
                           (long long)h->E * pitch4 <= (96LL << 20);
    if (h->cn_classes.size() == 1 && h->vn_classes.size() == 1 && coop_fits &&
        (h->cn_classes[0].degree < 6 || (h->use_pair && h->d_cn_pair != nullptr))) {
        if (h->coop_supported < 0) {
            int v = 0;
            CK(cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->device));
            h->coop_supported = v;
        }
        int smem = 0;
        CoopKernel k = h->coop_supported ? coop_kernel_for(h->cn_classes[0].degree, h->vn_classes[0].degree, early != 0, T,
                                                           h->match, &smem)
                                         : nullptr;
        if (k) {
            int occ;
            if ((rc = occupancy_of(h, (const void*)k, smem, &occ, kCoopThreads))) return rc;
            IbArgs b = a;
            b.tiles = (int)((pitch4 + 255) / 256);
            b.tpc_log2 = b.tiles > 4 ? 3 : b.tiles > 2 ? 2 : b.tiles == 2 ? 1 : 0;
            const int nps = (kCoopThreads / 32) >> b.tpc_log2;
            const long long want = std::max((h->cn_classes[0].count + nps - 1) / nps, (h->vn_classes[0].count + nps - 1) / nps);
            const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)occ * h->sm_count));
            CoopArgs c{};
            c.cn_nodes = h->cn_classes[0].d_nodes; c.vn_nodes = h->vn_classes[0].d_nodes;
            c.n_cn = h->cn_classes[0].count; c.n_vn = h->vn_classes[0].count;
            c.cn8 = h->d_cn8; c.vn8 = h->d_vn8;
            c.mc8 = h->match ? h->d_mc8 : nullptr; c.mv8 = h->match ? h->d_mv8 : nullptr;
            c.cn_pair = h->d_cn_pair; c.DCmax = h->DC; c.DVmax = h->DV;
            void* params[] = {&b, &c};
            if ((rc = prof.begin(2))) return rc;
            CK(cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(kCoopThreads), params, (size_t)smem, st));
            h->last_launches++; h->last_grid = grid; h->last_smem = smem;
            if ((rc = prof.end())) return rc;
            return IBLDPC_OK;
        }
    }
    // ---- small batches of the irregular codes whose degree sets are instantiated (802.11n, DVB-S2)
    if ((h->cn_classes.size() > 1 || h->vn_classes.size() > 1) && h->cn_classes.size() <= (size_t)kCoopMaxClasses &&
        h->vn_classes.size() <= (size_t)kCoopMaxClasses && coop_fits && h->use_pair &&
        h->d_cn_pair != nullptr) {
        if (h->coop_supported < 0) {
            int v = 0;
            CK(cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, h->device));
            h->coop_supported = v;
        }
        CoopClasses c{};
        c.n_cn_cls = (int)h->cn_classes.size(); c.n_vn_cls = (int)h->vn_classes.size();
        for (int i = 0; i < c.n_cn_cls; ++i) {
            c.cn_deg[i] = h->cn_classes[i].degree; c.cn_cnt[i] = h->cn_classes[i].count; c.cn_nodes[i] = h->cn_classes[i].d_nodes;
        }
        for (int i = 0; i < c.n_vn_cls; ++i) {
            c.vn_deg[i] = h->vn_classes[i].degree; c.vn_cnt[i] = h->vn_classes[i].count; c.vn_nodes[i] = h->vn_classes[i].d_nodes;
        }
        int smem = 0;
        CoopMultiKernel k = h->coop_supported ? coop_multi_kernel_for(c.cn_deg, c.n_cn_cls, c.vn_deg, c.n_vn_cls, early != 0, T,
                                                                      h->match, &smem)
                                              : nullptr;
        if (k) {
            int occ;
            if ((rc = occupancy_of(h, (const void*)k, smem, &occ, kCoopThreads))) return rc;
            IbArgs b = a;
            b.tiles = (int)((pitch4 + 255) / 256);
            b.tpc_log2 = b.tiles > 4 ? 3 : b.tiles > 2 ? 2 : b.tiles == 2 ? 1 : 0;
            const int nps = (kCoopThreads / 32) >> b.tpc_log2;
            long long want = 1;
            for (auto& cl : h->cn_classes) want = std::max<long long>(want, (cl.count + nps - 1) / nps);
            for (auto& cl : h->vn_classes) want = std::max<long long>(want, (cl.count + nps - 1) / nps);
            const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)occ * h->sm_count));
            c.cn8 = h->d_cn8; c.vn8 = h->d_vn8;
            c.mc8 = h->match ? h->d_mc8 : nullptr; c.mv8 = h->match ? h->d_mv8 : nullptr;
            c.cn_pair = h->d_cn_pair; c.DCmax = h->DC; c.DVmax = h->DV;
            void* params[] = {&b, &c};
            if ((rc = prof.begin(2))) return rc;
            CK(cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(kCoopThreads), params, (size_t)smem, st));
            h->last_launches++; h->last_grid = grid; h->last_smem = smem;
            if ((rc = prof.end())) return rc;
            return IBLDPC_OK;
        }
    }
    // ---- one launch per phase over all degree classes (ib_phase_n4.cuh) where the degree sets are instantiated
    // (by default for the 802.11n sets at every batch size; for the other instantiated sets in the launch-bound middle range
    // above the cooperative kernel's lane mode, where one launch per phase beats one per class -- measured on B200,
    // profiles/r02_small_and_mid_batches.txt: (3,6) B=2048 2.15 -> 1.90 ms, DVB-S2 B=1024 11.6 -> 8.6 ms, break-even at 4096)
    if (h->phase && (h->phase_default || (!h->phase_off_midrange && B <= h->phase_mid_max_frames)))
        return decode_ib_phase(h, a, imax, early, st);
    // launch geometry of one degree class (plan_geometry): tiles per CTA, tile groups, CTAs per tile group
    auto plan_launch = [&](IbArgs& b, const void* fn, int smem, int threads, int vec, int n_nodes, int* tile_groups, int* grid) -> int {
        int occ;
        int r = occupancy_of(h, fn, smem, &occ, threads);
        if (r) return r;
        static const bool no_plan = getenv("IBLDPC_NO_PLAN") != nullptr;
        b.tiles = (int)((pitch4 + 128 * vec - 1) / (128 * vec));
        plan_geometry((long long)occ * h->sm_count, threads / 32, b.tiles, n_nodes, no_plan, &b.tpc_log2, tile_groups, grid);
        return IBLDPC_OK;
    };
    // words per lane of the variable-node kernels: 4 up to degree 6, except for batches that fit one 2-word tile
    // (B <= 512), where 4 words would leave half of every warp's lanes without frames
    auto vn_vec_of = [&](int d) { return h->vn_vec ? h->vn_vec : (d <= 6 && pitch4 > 256 ? 4 : 2); };
    auto launch_cn = [&](int it) -> int {
        IbArgs b = a;
        b.it = it; b.iter0 = (it < 0);
        const int blk = it + 1;   // table block: 0 = iteration-0 tables
        b.lut = h->d_cn8 + (size_t)blk * (h->DC - 2) * TT;
        b.match = h->match ? h->d_mc8 + (size_t)blk * h->DC * T : nullptr;
        int r = prof.begin(it < 0 ? 2 : 0);
        if (r) return r;
        for (auto& c : h->cn_classes) {
            b.nst = c.degree - 2;
            const bool explicit_match = h->match && b.nst == 0;   // folded into the last stage otherwise
            b.dmax_match = h->match ? c.degree : 0;
            // tail-pair variant (cn_word_n4_pair): one LDS.64 of a host-composed row replaces the two last
            // look-ups of D-2 outputs; 2 words per lane only
            const bool pair = h->use_pair && c.degree >= h->n4_pair_min_degree && c.degree >= 4 && h->d_cn_pair != nullptr;
            const int vec = 2;
            int tile_groups;
            int smem = n4_table_bytes(n4_cn_words(c.degree, explicit_match)) + stage_scratch_bytes(b.nst, T, b.dmax_match);
            b.xp_col = -1;
            if (pair) {
                const size_t ci = (size_t)(&c - &h->cn_classes[0]);
                b.pair = h->d_cn_pair + ((size_t)blk * h->cn_classes.size() + ci) * (size_t)TT * 8;
                b.xp_col = c.degree - 5;    // column stored as 4*x (-1 for degree 4: raw messages feed the row)
                smem += (int)kPairBytes;
            }
            NodeKernel k = pair ? cn_n4_pair_kernel(c.degree, early != 0) : cn_n4_kernel_v2(c.degree, explicit_match, early != 0);
            if (!k) return fail(IBLDPC_E_INVALID, "no packed check-node kernel for degree " + std::to_string(c.degree));
            int threads = pair ? cn_n4_pair_threads(c.degree) : kThreads;
            if (pair && h->cn_threads != 512 && cn_n4_pair_kernel_1024(c.degree, early != 0)) {
                k = cn_n4_pair_kernel_1024(c.degree, early != 0);
                threads = 1024;
            }
            IbArgs bt = b;
            if (pair && c.degree >= 6 && c.degree <= h->cn_tri_max_degree && h->d_cn3 != nullptr && h->cn_threads == 0 &&
                h->n4_pair_min_degree <= 6) {
                // the first two stages of every chain through the three-input table (degree 6: 7 + 1 look-ups per frame
                // instead of 10 + 1, degree 7: 12 + 1 instead of 15 + 1, degree 8: 18 + 1 instead of 21 + 1; 192 KB of
                // shared memory per CTA).  Degrees 7 and 8 stage the columns 2..D-3 only, as local columns 0..
                bt.lut_all = h->d_cn3 + (size_t)blk * kTS * kTS * kTS;
                if (c.degree >= 7) {
                    bt.lut = b.lut + (size_t)2 * TT;
                    bt.nst = c.degree - 4;
                    bt.xp_col = c.degree - 7;
                }
                k = cn_n4_tri_kernel(c.degree, early != 0);
                threads = 1024;
                smem = (int)kPairBytes + kTripleBytes + n4_table_bytes(1) + stage_scratch_bytes(bt.nst, T, bt.dmax_match);
            }
            int grid;
            if ((r = plan_launch(bt, (const void*)k, smem, threads, vec, c.count, &tile_groups, &grid))) return r;
            k<<<dim3(grid, tile_groups), threads, smem, st>>>(bt, c.d_nodes, c.count);
            h->last_launches++; h->last_grid = grid * tile_groups; h->last_smem = smem;
        }
        return prof.end();
    };
    auto launch_vn = [&](int it, bool decide) -> int {
        IbArgs b = a;
        b.it = it; b.iter0 = 0;
        if (decide) {
            b.lut = h->d_vn8; b.vn_it_stride = (long long)h->DV * TT; b.match = nullptr;
        } else {
            b.lut = h->d_vn8 + (size_t)it * h->DV * TT;
            b.match = h->match ? h->d_mv8 + (size_t)it * h->DV * T : nullptr;
        }
        int r = prof.begin(decide ? 2 : 1);
        if (r) return r;
        for (auto& c : h->vn_classes) {
            b.nst = decide ? c.degree : c.degree - 1;
            b.dmax_match = b.match != nullptr ? c.degree : 0;
            // tail-pair variant (vn_word_n4_pair) for the high degrees: one LDS.64 of a host-composed row replaces
            // the two last look-ups of D-2 outputs; 2 words per lane, 512-thread CTAs share one table set
            const bool pair = !decide && h->use_pair && h->d_vn_pair != nullptr && c.degree >= h->vn_pair_min_degree &&
                              c.degree >= 3 && (h->vn_vec == 0 || h->vn_vec == 2);
            const int vec = pair ? 2 : vn_vec_of(c.degree);
            int vec_used = vec;
            int tile_groups;
            int smem = n4_table_bytes(n4_vn_words(c.degree, decide)) + stage_scratch_bytes(b.nst, T, b.dmax_match);
            int threads = kThreads;
            b.xp_col = -1;
            NodeKernel k;
            const bool triple = !decide && !pair && c.degree == 3 && h->d_vn3 != nullptr && h->vn_vec == 0 && h->vn_threads == 0;
            IbArgs b3 = b;
            if (triple) {
                // degree 3: three look-ups per frame in the three-input table F (128 KB per CTA) instead of five
                b3.lut = h->d_vn3 + (size_t)it * kTS * kTS * kTS;
                smem = kTripleBytes;
                threads = 1024;
                vec_used = pitch4 > 256 ? 4 : 2;
                k = vn3_n4_kernel(vec_used);
            } else if (pair) {
                const size_t ci = (size_t)(&c - &h->vn_classes[0]);
                b.pair = h->d_vn_pair + ((size_t)it * h->vn_classes.size() + ci) * (size_t)TT * 8;
                b.xp_col = c.degree - 4;    // column stored as 4*x (-1 for degree 3: the channel value feeds the row)
                smem += (int)kPairBytes;
                // measured on B200: degrees 8-9 spill at the 64 registers a 2 x 512-thread residency allows (DVB-S2 d_v=8:
                // 0.387 ms with 256 threads vs 0.420 ms), degree >= 10 has one table set per SM either way (WLAN d_v=11:
                // 0.195 ms with 512 threads vs 0.229 ms)
                // (second half of round 2: 768-thread CTAs for degrees 8-9 -- one CTA per SM, 24 warps at 80 registers, no spills)
                threads = h->vn_pair_threads ? h->vn_pair_threads : (c.degree == 8 || c.degree == 9) ? 768 : 512;
                if (threads == 768 && !vn_n4_pair_kernel(c.degree, 768)) threads = 512;
                k = vn_n4_pair_kernel(c.degree, threads);
            } else {
                k = vec == 4 ? vn_n4_kernel_v4(c.degree, decide) : vn_n4_kernel_v2(c.degree, decide);
                if (!k && vec == 4) {   // 4-word kernels exist up to degree 6 only
                    vec_used = 2;
                    k = vn_n4_kernel_v2(c.degree, decide);
                }
                if (!decide && vec_used == 4 && h->vn_threads != 256 && vn_n4_kernel_v4_1024(c.degree)) {
                    k = vn_n4_kernel_v4_1024(c.degree);
                    threads = 1024;
                }
            }
            if (!k) return fail(IBLDPC_E_INVALID, "no packed variable-node kernel for degree " + std::to_string(c.degree));
            int grid;
            IbArgs& bl = triple ? b3 : b;
            if ((r = plan_launch(bl, (const void*)k, smem, threads, vec_used, c.count, &tile_groups, &grid))) return r;
            k<<<dim3(grid, tile_groups), threads, smem, st>>>(bl, c.d_nodes, c.count);
            h->last_launches++;
        }
        return prof.end();
    };
    if ((rc = launch_cn(-1))) return rc;
    for (int it = 0; it < imax - 1; ++it) {
        if ((rc = launch_vn(it, false))) return rc;
        if ((rc = launch_cn(it))) return rc;
    }
    if ((rc = launch_vn(0, true))) return rc;
    CK(cudaGetLastError());
    return IBLDPC_OK;
}

// i_num and the input-range flag of the last decode on workspace `w` (synchronises `st`)
int read_back_status(Workspace& w, cudaStream_t st, int32_t* i_num_host)
{
    int host2[2] = {0, 0};
    CK(cudaMemcpyAsync(&host2[0], w.inum, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&host2[1], w.flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (i_num_host) *i_num_host = host2[0];
    if (host2[1] != 0) CK(cudaMemsetAsync(w.flags, 0, sizeof(int), st));   // reported once
    if (host2[1] != 0)
        return fail(IBLDPC_E_INVALID, "channel cluster indices must lie in [0, cardinality_T_channel): the decode clamped out-of-range values");
    return IBLDPC_OK;
}

int ensure_n4_buffers(ibldpc_decoder* h, Workspace& w, long long B)
{
    const long long pitch4 = ((B + 1) / 2 + 15) / 16 * 16;
    void* p = w.msg;
    int rc = ensure(&p, &w.msg_bytes, (size_t)h->E * (size_t)pitch4);
    w.msg = (uint8_t*)p;
    if (rc) return rc;
    p = w.ch4;
    rc = ensure(&p, &w.ch4_bytes, (size_t)h->N * (size_t)pitch4);
    w.ch4 = (uint8_t*)p;
    return rc;
}

int check_decode_args(ibldpc_decoder* h, int64_t B, int imax)
{
    if (!h) return fail(IBLDPC_E_INVALID, "null handle");
    if (!h->have_luts) return fail(IBLDPC_E_STATE, "ibldpc_set_luts must be called before decoding");
    if (B <= 0) return fail(IBLDPC_E_INVALID, "B must be positive");
    if (B > 0x7fffffffLL - 1024) return fail(IBLDPC_E_INVALID, "B too large");
    if (imax < 1 || imax > h->lut_imax)
        return fail(IBLDPC_E_INVALID, "imax must be in [1, " + std::to_string(h->lut_imax) + "] (the tables hold that many iterations)");
    if (imax >= kMaxIter) return fail(IBLDPC_E_INVALID, "imax too large");
    return IBLDPC_OK;
}

template <typename T>
int launch_pad(const T* src, T* dst, long long rows, long long B, long long pitch, T fill, cudaStream_t st)
{
    const long long n = rows * pitch;
    int grid = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    pad_rows_kernel<T><<<grid, 256, 0, st>>>(src, dst, rows, B, pitch, fill);
    CK(cudaGetLastError());
    return IBLDPC_OK;
}
template <typename T>
int launch_unpad(const T* src, T* dst, long long rows, long long B, long long pitch, cudaStream_t st)
{
    const long long n = rows * B;
    int grid = (int)std::min<long long>((n + 255) / 256, 148 * 16);
    unpad_rows_kernel<T><<<grid, 256, 0, st>>>(src, dst, rows, B, pitch);
    CK(cudaGetLastError());
    return IBLDPC_OK;
}

int ensure_pad(Workspace& w, size_t bytes)
{
    if (w.pad_bytes >= bytes && w.padbuf_in) return IBLDPC_OK;
    if (w.padbuf_in) CK(cudaFree(w.padbuf_in));
    if (w.padbuf_out) CK(cudaFree(w.padbuf_out));
    w.padbuf_in = w.padbuf_out = nullptr;
    w.pad_bytes = 0;
    if (cudaMalloc((void**)&w.padbuf_in, bytes) != cudaSuccess || cudaMalloc((void**)&w.padbuf_out, bytes) != cudaSuccess)
        return fail(IBLDPC_E_NOMEM, "cudaMalloc of padding buffers failed");
    w.pad_bytes = bytes;
    return IBLDPC_OK;
}

// ------------------------------------------------------------------------------------------
// LLR decoders
// ------------------------------------------------------------------------------------------
template <typename F, int ALGO>
int decode_llr_padded(ibldpc_decoder* h, Workspace& w, const F* ch, long long pitch, long long B, int imax, int early,
                      F* out, cudaStream_t st)
{
    constexpr int V = VecOf<F>::N;
    int rc = ensure_ws_common(h, w);
    if (rc) return rc;
    const size_t need = (size_t)h->E * (size_t)pitch * sizeof(F);
    if (w.llr_bytes < need || !w.cin) {
        if (w.cin) CK(cudaFree(w.cin));
        if (w.vin) CK(cudaFree(w.vin));
        w.cin = w.vin = nullptr;
        w.llr_bytes = 0;
        if (cudaMalloc(&w.cin, need) != cudaSuccess || cudaMalloc(&w.vin, need) != cudaSuccess)
            return fail(IBLDPC_E_NOMEM, "cudaMalloc of LLR message arrays failed");
        w.llr_bytes = need;
    }
    CK(cudaMemsetAsync(w.flags + 1, 0, sizeof(int) * (size_t)std::max(imax, 1), st));   // flags[0] (input range) is sticky until read
    LlrArgs a{};
    a.sc = h->d_sc; a.deg_c = h->d_dc; a.tc = h->d_tc; a.sv = h->d_sv; a.deg_v = h->d_dv; a.tv = h->d_tv;
    a.n_var = h->N; a.n_chk = h->M;
    a.ch = ch; a.cin = w.cin; a.vin = w.vin; a.out = out;
    a.pitch = pitch; a.B = (int)B; a.tiles = (int)((pitch + 32 * V - 1) / (32 * V));
    a.flags = w.flags + 1; a.inum = w.inum; a.early = early; a.imax = imax;
    h->last_launches = 0;
    auto run_vn = [&](int mode, int it) -> int {
        LlrArgs b = a;
        b.it = it;
        for (auto& c : h->vn_classes) {
            LlrNodeKernel k = llr_vn_kernel_for(sizeof(F) == 8, mode, c.degree);
            int grid;
            int r = grid_for(h, (const void*)k, 0, (long long)c.count * b.tiles, &grid);
            if (r) return r;
            k<<<grid, kThreads, 0, st>>>(b, c.d_nodes, c.count);
            h->last_launches++;
        }
        return IBLDPC_OK;
    };
    // low-batch min-sum: lanes = edges (warp-shuffle min/argmin), see llr_cn_minsum_shfl_kernel
    const bool low_batch = (ALGO == 0) && (B <= 8) && (h->dc_max <= 32) && !getenv("IBLDPC_NO_SHFL");
    int seg_log2 = 1;
    while ((1 << seg_log2) < h->dc_max) ++seg_log2;
    auto run_cn = [&](int it) -> int {
        LlrArgs b = a;
        b.it = it;
        if (low_batch) {
            LlrShflKernel k = llr_cn_minsum_shfl_kernel_for(sizeof(F) == 8);
            const long long per_warp = 32 >> seg_log2;
            int grid;
            int r = grid_for(h, (const void*)k, 0, ((long long)h->M + per_warp - 1) / per_warp, &grid);
            if (r) return r;
            k<<<grid, kThreads, 0, st>>>(b, seg_log2);
            h->last_launches++;
            return IBLDPC_OK;
        }
        for (auto& c : h->cn_classes) {
            // float64 BP: forward/backward recursion (ALGO 2) unless the reference's operation order is asked for
            const bool bp_sequential = getenv("IBLDPC_BP_SEQUENTIAL") != nullptr;
            // float64 BP: forward/backward recursion in the likelihood-ratio domain (3); IBLDPC_BP_LOGDOMAIN=1: the same
            // recursion with the reference box-plus expression per operation (2); IBLDPC_BP_SEQUENTIAL=1: reference order (1)
            const bool bp_logdomain = getenv("IBLDPC_BP_LOGDOMAIN") != nullptr;
            const int algo_k = (ALGO == 1 && sizeof(F) == 8 && !bp_sequential) ? (bp_logdomain ? 2 : 3) : ALGO;
            LlrNodeKernel k = llr_cn_kernel_for(sizeof(F) == 8, algo_k, c.degree);
            int grid;
            int r = grid_for(h, (const void*)k, 0, (long long)c.count * b.tiles, &grid);
            if (r) return r;
            k<<<grid, kThreads, 0, st>>>(b, c.d_nodes, c.count);
            h->last_launches++;
        }
        return IBLDPC_OK;
    };
    if ((rc = run_vn(2, 0))) return rc;
    for (int it = 0; it < imax - 1; ++it) {
        if ((rc = run_cn(it))) return rc;
        if ((rc = run_vn(0, it))) return rc;
        if (early) {
            LlrArgs b = a;
            b.it = it;
            int grid;
            LlrSynKernel ks = llr_syndrome_kernel_for(sizeof(F) == 8);
            if ((rc = grid_for(h, (const void*)ks, 0, (long long)h->M * b.tiles, &grid))) return rc;
            ks<<<grid, kThreads, 0, st>>>(b);
            h->last_launches++;
        }
    }
    if ((rc = run_vn(1, 0))) return rc;
    CK(cudaGetLastError());
    return IBLDPC_OK;
}

template <typename F>
int decode_llr_typed(ibldpc_decoder* h, int algo, const F* ch, int64_t B, int imax, int early, F* out,
                     int32_t* i_num_host, cudaStream_t st)
{
    constexpr int V = VecOf<F>::N;
    Workspace& w = h->ws[0];
    const long long pitch = (B + V - 1) / V * V;
    const bool aligned = (B % V == 0) && ((uintptr_t)ch % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const F* chp = ch;
    F* outp = out;
    int rc;
    if (!aligned) {
        if ((rc = ensure_pad(w, (size_t)h->N * pitch * sizeof(F)))) return rc;
        if ((rc = launch_pad<F>(ch, (F*)w.padbuf_in, h->N, B, pitch, F(0), st))) return rc;
        chp = (const F*)w.padbuf_in;
        outp = (F*)w.padbuf_out;
    }
    rc = algo == IBLDPC_ALGO_MINSUM ? decode_llr_padded<F, 0>(h, w, chp, pitch, B, imax, early, outp, st)
                                    : decode_llr_padded<F, 1>(h, w, chp, pitch, B, imax, early, outp, st);
    if (rc) return rc;
    if (!aligned && (rc = launch_unpad<F>(outp, out, h->N, B, pitch, st))) return rc;
    h->last_stream = st;
    h->last_ws = 0;
    if (i_num_host) {
        CK(cudaMemcpyAsync(i_num_host, w.inum, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return IBLDPC_OK;
}

// layered schedule (llr_layered.cu): same buffer handling as decode_llr_typed
template <typename F>
int decode_llr_layered_typed(ibldpc_decoder* h, int algo, const F* ch, int64_t B, int imax, int early, F* out,
                             int32_t* i_num_host, cudaStream_t st)
{
    constexpr int V = VecOf<F>::N;
    Workspace& w = h->ws[0];
    int rc = ensure_ws_common(h, w);
    if (rc) return rc;
    if ((rc = layered_prepare(h))) return rc;
    const long long pitch = (B + V - 1) / V * V;
    const bool aligned = (B % V == 0) && ((uintptr_t)ch % 16 == 0) && ((uintptr_t)out % 16 == 0);
    const F* chp = ch;
    F* outp = out;
    if (!aligned) {
        if ((rc = ensure_pad(w, (size_t)h->N * pitch * sizeof(F)))) return rc;
        if ((rc = launch_pad<F>(ch, (F*)w.padbuf_in, h->N, B, pitch, F(0), st))) return rc;
        chp = (const F*)w.padbuf_in;
        outp = (F*)w.padbuf_out;
    }
    if constexpr (sizeof(F) == 4) rc = decode_llr_layered_f32(h, w, algo, chp, pitch, B, imax, early, outp, st);
    else rc = decode_llr_layered_f64(h, w, algo, chp, pitch, B, imax, early, outp, st);
    if (rc) return rc;
    if (!aligned && (rc = launch_unpad<F>(outp, out, h->N, B, pitch, st))) return rc;
    h->last_stream = st;
    h->last_ws = 0;
    if (i_num_host) {
        CK(cudaMemcpyAsync(i_num_host, w.inum, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    return IBLDPC_OK;
}

template <int SRC>
int quantize_common(int device, const double* x_dev, int64_t n, const double* limits_host, int card,
                    const double* llr_host, int out_kind, uint64_t seed, uint64_t offset, void* out_dev, void* stream)
{
    if (n < 0 || card < 1 || card > 4096) return fail(IBLDPC_E_INVALID, "bad n / card");
    if (out_kind != 3 && !limits_host) return fail(IBLDPC_E_INVALID, "limits missing");
    if ((out_kind == 1 || out_kind == 2) && !llr_host) return fail(IBLDPC_E_INVALID, "LLR vector missing");
    if (n == 0) return IBLDPC_OK;
    DeviceGuard guard_(device);
    CK(guard_.err);
    cudaStream_t st = (cudaStream_t)stream;
    double* d_tab = nullptr;
    CK(cudaMallocAsync((void**)&d_tab, sizeof(double) * 2 * card, st));
    if (limits_host) CK(cudaMemcpyAsync(d_tab, limits_host, sizeof(double) * card, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(d_tab, 0, sizeof(double) * card, st));
    if (llr_host) CK(cudaMemcpyAsync(d_tab + card, llr_host, sizeof(double) * card, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(d_tab + card, 0, sizeof(double) * card, st));
    // pageable host memory: the copies above complete (are staged) before the call returns
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    const size_t smem = sizeof(double) * 2 * card;
    switch (out_kind) {
    case 0: quantize_kernel<SRC, 0><<<grid, 256, smem, st>>>(x_dev, n, d_tab, card, d_tab + card, seed, offset, out_dev); break;
    case 1: quantize_kernel<SRC, 1><<<grid, 256, smem, st>>>(x_dev, n, d_tab, card, d_tab + card, seed, offset, out_dev); break;
    case 2: quantize_kernel<SRC, 2><<<grid, 256, smem, st>>>(x_dev, n, d_tab, card, d_tab + card, seed, offset, out_dev); break;
    default: quantize_kernel<SRC, 3><<<grid, 256, smem, st>>>(x_dev, n, d_tab, card, d_tab + card, seed, offset, out_dev); break;
    }
    CK(cudaGetLastError());
    CK(cudaFreeAsync(d_tab, st));
    return IBLDPC_OK;
}

template <typename E>
int count_errors_common(int device, const E* out_dev, int64_t rows, int64_t B, E threshold, const uint8_t* ref_bits,
                        int64_t* counters_host, int64_t* counters_dev, void* stream)
{
    if (!out_dev || (!counters_host && !counters_dev) || rows < 0 || B <= 0) return fail(IBLDPC_E_INVALID, "bad arguments");
    DeviceGuard guard_(device);
    CK(guard_.err);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d_cnt = (unsigned long long*)counters_dev;   // async variant: accumulate in place
    int* d_fe = nullptr;
    if (!counters_dev) {
        CK(cudaMallocAsync((void**)&d_cnt, sizeof(unsigned long long) * 2, st));
        CK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * 2, st));
    }
    CK(cudaMallocAsync((void**)&d_fe, sizeof(int) * B, st));
    CK(cudaMemsetAsync(d_fe, 0, sizeof(int) * B, st));
    if (rows > 0) {
        dim3 grid((unsigned)((B + 255) / 256), (unsigned)((rows + 63) / 64));
        count_errors_kernel<E><<<grid, 256, 0, st>>>(out_dev, rows, B, threshold, ref_bits, d_cnt, d_fe);
        count_frames_kernel<<<(unsigned)std::min<int64_t>((B + 255) / 256, 1024), 256, 0, st>>>(d_fe, B, d_cnt);
    }
    CK(cudaGetLastError());
    CK(cudaFreeAsync(d_fe, st));
    if (!counters_dev) {
        unsigned long long hc[2];
        CK(cudaMemcpyAsync(hc, d_cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaFreeAsync(d_cnt, st));
        counters_host[0] = (int64_t)hc[0];
        counters_host[1] = (int64_t)hc[1];
    }
    return IBLDPC_OK;
}

}  // namespace

namespace {
template <typename Fn>
void parallel_rows(int64_t rows, int nthreads, Fn fn)
{
    nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, rows));
    if (nthreads == 1) { fn(0, rows); return; }
    std::vector<std::thread> pool;
    const int64_t per = (rows + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t lo = t * per, hi = std::min<int64_t>(rows, lo + per);
        if (lo < hi) pool.emplace_back([=] { fn(lo, hi); });
    }
    for (auto& th : pool) th.join();
}
}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* ibldpc_last_error(void) { return g_err.c_str(); }

int ibldpc_create(const ibldpc_code_desc* code, int device, ibldpc_handle* out)
{
    if (!code || !out) return fail(IBLDPC_E_INVALID, "null argument");
    const int N = code->n_var, M = code->n_chk, E = code->n_edge;
    if (N <= 0 || M <= 0 || E <= 0) return fail(IBLDPC_E_INVALID, "empty graph");
    if (!code->inbox_start_chk || !code->degree_chk || !code->target_cells_chk || !code->inbox_start_var ||
        !code->degree_var || !code->target_cells_var)
        return fail(IBLDPC_E_INVALID, "null table pointer");
    // ---- validation: prefix sums, inverse permutations
    long long acc = 0;
    for (int c = 0; c < M; ++c) {
        if (code->inbox_start_chk[c] != acc) return fail(IBLDPC_E_INVALID, "inbox_start_chk is not the exclusive prefix sum of degree_chk");
        if (code->degree_chk[c] < 2 || code->degree_chk[c] > kMaxGenericDeg)
            return fail(IBLDPC_E_INVALID, "check-node degrees must lie in [2, 64]");
        acc += code->degree_chk[c];
    }
    if (acc != E) return fail(IBLDPC_E_INVALID, "sum(degree_chk) != n_edge");
    acc = 0;
    for (int v = 0; v < N; ++v) {
        if (code->inbox_start_var[v] != acc) return fail(IBLDPC_E_INVALID, "inbox_start_var is not the exclusive prefix sum of degree_var");
        if (code->degree_var[v] < 1 || code->degree_var[v] > kMaxGenericDeg)
            return fail(IBLDPC_E_INVALID, "variable-node degrees must lie in [1, 64]");
        acc += code->degree_var[v];
    }
    if (acc != E) return fail(IBLDPC_E_INVALID, "sum(degree_var) != n_edge");
    for (int e = 0; e < E; ++e) {
        const int t = code->target_cells_chk[e];
        if (t < 0 || t >= E || code->target_cells_var[t] != e)
            return fail(IBLDPC_E_INVALID, "target_cells_chk / target_cells_var are not inverse permutations");
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(IBLDPC_E_INVALID, "no such CUDA device");
    DeviceGuard guard_(device);
    CK(guard_.err);
    ibldpc_decoder* h = new ibldpc_decoder();
    h->device = device;
    h->N = N; h->M = M; h->E = E;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    std::vector<int> dc(code->degree_chk, code->degree_chk + M), dv(code->degree_var, code->degree_var + N);
    h->dc_max = *std::max_element(dc.begin(), dc.end());
    h->dc_min = *std::min_element(dc.begin(), dc.end());
    h->dv_max = *std::max_element(dv.begin(), dv.end());
    h->dv_min = *std::min_element(dv.begin(), dv.end());
    h->h_sc.assign(code->inbox_start_chk, code->inbox_start_chk + M);
    h->h_sv.assign(code->inbox_start_var, code->inbox_start_var + N);
    h->h_dc = dc;
    h->h_dv = dv;
    // variable index of every CN-major row: row tv[sv[v]+k] belongs to variable v
    std::vector<int> vidx(E);
    for (int v = 0; v < N; ++v)
        for (int k = 0; k < dv[v]; ++k) vidx[code->target_cells_var[code->inbox_start_var[v] + k]] = v;
    int rc = 0;
    rc |= upload(&h->d_sc, code->inbox_start_chk, M);
    rc |= upload(&h->d_dc, code->degree_chk, M);
    rc |= upload(&h->d_tc, code->target_cells_chk, E);
    rc |= upload(&h->d_sv, code->inbox_start_var, N);
    rc |= upload(&h->d_dv, code->degree_var, N);
    rc |= upload(&h->d_tv, code->target_cells_var, E);
    rc |= upload(&h->d_vidx, vidx.data(), E);
    h->h_vidx = vidx;
    if (!rc) rc = build_classes(dc, h->cn_classes);
    if (!rc) rc = build_classes(dv, h->vn_classes);
    if (rc) { ibldpc_destroy(h); return rc; }
    // the LLR decoders need no tables
    h->have_luts = false;
    *out = h;
    return IBLDPC_OK;
}

int ibldpc_set_luts(ibldpc_handle h, const ibldpc_lut_desc* L)
{
    if (!h || !L) return fail(IBLDPC_E_INVALID, "null argument");
    const int T = L->card_decoder, Tc = L->card_channel, imax = L->imax, DC = L->cn_degree, DV = L->vn_degree;
    if (T < 2 || T > 256 || Tc < 2 || Tc > 256) return fail(IBLDPC_E_INVALID, "cardinalities must lie in [2, 256]");
    if (imax < 1 || imax >= kMaxIter) return fail(IBLDPC_E_INVALID, "imax out of range");
    if (DC < h->dc_max || DV < h->dv_max)
        return fail(IBLDPC_E_INVALID, "cn_degree / vn_degree must be at least the maximum node degrees of the code");
    if (!L->cn_lut || !L->vn_lut) return fail(IBLDPC_E_INVALID, "null LUT pointer");
    const long long need_cn = (long long)Tc * Tc + (long long)(DC - 3) * Tc * T + (long long)(imax - 1) * (DC - 2) * T * T;
    const long long need_vn = (long long)imax * ((long long)Tc * T + (long long)(DV - 1) * T * T);
    if (L->cn_lut_len < need_cn)
        return fail(IBLDPC_E_INVALID, "Trellis_checknode_vector_a too short: need " + std::to_string(need_cn) + ", got " + std::to_string(L->cn_lut_len));
    if (L->vn_lut_len < need_vn)
        return fail(IBLDPC_E_INVALID, "Trellis_varnode_vector_a too short: need " + std::to_string(need_vn) + ", got " + std::to_string(L->vn_lut_len));
    const bool match = L->cn_match != nullptr || L->vn_match != nullptr;
    if (match) {
        if (!L->cn_match || !L->vn_match) return fail(IBLDPC_E_INVALID, "both matching vectors are required");
        if (L->cn_match_len < (long long)imax * DC * T || L->vn_match_len < (long long)imax * DV * T)
            return fail(IBLDPC_E_INVALID, "matching vectors too short (need imax*degree*T entries)");
    }
    auto to_u8 = [&](const int32_t* src, long long n, std::vector<uint8_t>& dst, const char* name) -> int {
        dst.resize((size_t)std::max<long long>(n, 1));
        for (long long i = 0; i < n; ++i) {
            if (src[i] < 0 || src[i] >= T) return fail(IBLDPC_E_INVALID, std::string(name) + ": entry out of range [0,T)");
            dst[(size_t)i] = (uint8_t)src[i];
        }
        return IBLDPC_OK;
    };
    std::vector<uint8_t> cn, vn, mc, mv;
    int rc;
    if ((rc = to_u8(L->cn_lut, std::max<long long>(need_cn, 0), cn, "Trellis_checknode_vector_a"))) return rc;
    if ((rc = to_u8(L->vn_lut, need_vn, vn, "Trellis_varnode_vector_a"))) return rc;
    if (match) {
        if ((rc = to_u8(L->cn_match, (long long)imax * DC * T, mc, "matching_vector_checknode"))) return rc;
        if ((rc = to_u8(L->vn_match, (long long)imax * DV * T, mv, "matching_vector_varnode"))) return rc;
    }
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    CK(cudaDeviceSynchronize());
    // from here on the old tables are gone: a failure below must not leave a decodable handle behind
    h->have_luts = false;
    h->fast = h->nib = h->t32 = false;
    for (uint8_t** p : {&h->d_cn8, &h->d_vn8, &h->d_mc8, &h->d_mv8}) {
        if (*p) CK(cudaFree(*p));
        *p = nullptr;
    }
    if ((rc = upload(&h->d_cn8, cn.data(), cn.size()))) return rc;
    if ((rc = upload(&h->d_vn8, vn.data(), vn.size()))) return rc;
    if (match) {
        if ((rc = upload(&h->d_mc8, mc.data(), mc.size()))) return rc;
        if ((rc = upload(&h->d_mv8, mv.data(), mv.size()))) return rc;
    }
    h->h_cn8 = cn; h->h_vn8 = vn; h->h_mc8 = mc; h->h_mv8 = mv;
    h->T = T; h->Tc = Tc; h->lut_imax = imax; h->DC = DC; h->DV = DV; h->match = match;
    h->tshift = -1;
    for (int s = 1; s <= 8; ++s)
        if ((1 << s) == T) h->tshift = s - 1;
    // ---- fast path: lane-striped shared-memory tables
    auto words = [](int cols) { return std::max(1, (cols + 3) / 4); };
    h->Wc = words((DC - 2) + (match ? 1 : 0));
    h->Wv = words((DV - 1) + (match ? 1 : 0));
    h->Wo = words(DV);
    h->nrows_c = std::max(T * T, match ? DC * T : 0);
    h->nrows_v = std::max(T * T, match ? DV * T : 0);
    h->nrows_o = T * T;
    const size_t smem_max = 227 * 1024 - 4096;   // leaves room for the staging scratch
    h->fast = (T <= 16) && (Tc == T) && h->dc_max <= kMaxFastDc && h->dv_max <= kMaxFastDv &&
              (size_t)h->nrows_c * h->Wc * 128 <= smem_max && (size_t)h->nrows_v * h->Wv * 128 <= smem_max &&
              (size_t)h->nrows_o * h->Wo * 128 <= smem_max;
    if (getenv("IBLDPC_FORCE_GENERIC")) h->fast = false;
    // 16 < |T| <= 32 (the reference's 802.11n design uses 32): byte messages, per-lane replicated 32 x 32 stage tables
    // (ib_kernels_t32.cuh).  Degree-2 checks (explicit matching column) and T_c != T stay on the generic path.
    h->t32 = !h->fast && T > 16 && T <= 32 && Tc == T && h->dc_min >= 3 && h->dc_max <= 10 && h->dv_max <= 12 &&
             getenv("IBLDPC_FORCE_GENERIC") == nullptr && getenv("IBLDPC_NO_T32") == nullptr;
    // packed-nibble messages need an even |T| (address terms are shifts of the packed word)
    h->nib = h->fast && (T % 2 == 0) && getenv("IBLDPC_NO_NIBBLE") == nullptr;
    h->vn_vec = 0;
    if (const char* e = getenv("IBLDPC_VN_VEC")) h->vn_vec = atoi(e) == 4 ? 4 : atoi(e) == 2 ? 2 : 0;
    h->use_pair = getenv("IBLDPC_NO_PAIR") == nullptr;
    if (const char* pm = getenv("IBLDPC_PAIR_MIN_DEGREE")) h->pair_min_degree = h->n4_pair_min_degree = std::max(4, atoi(pm));
    if (h->d_cn_pair) { CK(cudaFree(h->d_cn_pair)); h->d_cn_pair = nullptr; }
    if (h->fast && h->use_pair) {
        // Composed tail-pair tables of the check-node kernel (cn_word_pair): for every iteration block and
        // every check degree d >= 4, G(a,b)[x] = S_b'( S_a(x, a), b ) with S_a = stage d-4, S_b = stage d-3
        // (matching folded into S_b), 16 nibbles per (a,b) row.
        const int TT = T * T;
        const size_t ncls = h->cn_classes.size();
        std::vector<uint8_t> pair((size_t)imax * ncls * TT * 8, 0);
        for (int blk = 0; blk < imax; ++blk)
            for (size_t ci = 0; ci < ncls; ++ci) {
                const int d = h->cn_classes[ci].degree;
                if (d < 4) continue;
                const uint8_t* Sa = cn.data() + ((size_t)blk * (DC - 2) + (d - 4)) * TT;
                const uint8_t* Sb = cn.data() + ((size_t)blk * (DC - 2) + (d - 3)) * TT;
                const uint8_t* mrow = match ? mc.data() + ((size_t)blk * DC + (d - 1)) * T : nullptr;
                uint8_t* dst = pair.data() + ((size_t)blk * ncls + ci) * TT * 8;
                for (int a = 0; a < T; ++a)
                    for (int b = 0; b < T; ++b)
                        for (int x = 0; x < T; ++x) {
                            int v = Sb[Sa[x * T + a] * T + b];
                            if (mrow) v = mrow[v];
                            dst[(a * T + b) * 8 + (x >> 1)] |= (uint8_t)(v << ((x & 1) * 4));
                        }
            }
        if ((rc = upload(&h->d_cn_pair, pair.data(), pair.size()))) return rc;
        h->h_cn_pair.swap(pair);
    }
    if (h->d_vn_pair) { CK(cudaFree(h->d_vn_pair)); h->d_vn_pair = nullptr; }
    if (const char* e = getenv("IBLDPC_VN_PAIR_MIN_DEGREE")) h->vn_pair_min_degree = std::max(3, atoi(e));
    h->cn_threads = getenv("IBLDPC_CN_THREADS") ? atoi(getenv("IBLDPC_CN_THREADS")) : 0;
    h->vn_threads = getenv("IBLDPC_VN_THREADS") ? atoi(getenv("IBLDPC_VN_THREADS")) : 0;
    if (const char* e = getenv("IBLDPC_VN_PAIR_THREADS")) h->vn_pair_threads = atoi(e) == 768 ? 768 : atoi(e) == 512 ? 512 : atoi(e) == 256 ? 256 : 0;
    if (h->nib && h->use_pair) {
        // Composed tail-pair tables of the variable-node update (vn_word_n4_pair): for every iteration and every
        // variable-node degree d >= 3, G(a,b)[x] = M_d( S_{d-2}( S_{d-3}(x, a), b ) ), 16 nibbles per (a,b) row
        // (stage 0 is indexed by the channel value; with Tc == T every stage block holds T*T entries).
        const int TT = T * T;
        const size_t ncls = h->vn_classes.size();
        std::vector<uint8_t> vpair((size_t)imax * ncls * TT * 8, 0);
        for (int it = 0; it < imax; ++it)
            for (size_t ci = 0; ci < ncls; ++ci) {
                const int d = h->vn_classes[ci].degree;
                if (d < 3) continue;
                const uint8_t* Sa = vn.data() + ((size_t)it * DV + (d - 3)) * TT;
                const uint8_t* Sb = vn.data() + ((size_t)it * DV + (d - 2)) * TT;
                const uint8_t* mrow = match ? mv.data() + ((size_t)it * DV + (d - 1)) * T : nullptr;
                uint8_t* dst = vpair.data() + ((size_t)it * ncls + ci) * TT * 8;
                for (int a = 0; a < T; ++a)
                    for (int b = 0; b < T; ++b)
                        for (int x = 0; x < T; ++x) {
                            int v = Sb[Sa[x * T + a] * T + b];
                            if (mrow) v = mrow[v];
                            dst[(a * T + b) * 8 + (x >> 1)] |= (uint8_t)(v << ((x & 1) * 4));
                        }
            }
        if ((rc = upload(&h->d_vn_pair, vpair.data(), vpair.size()))) return rc;
        h->h_vn_pair.swap(vpair);
    }
    // Three-input table of the degree-3 variable-node update (ib_triple_n4.cuh): F(a, b, c) = M_3( S_1( S_0(a, b), c ) ),
    // index b*256 + c*16 + a (channel value minor), one byte per entry, for every iteration
    if (h->d_vn3) { CK(cudaFree(h->d_vn3)); h->d_vn3 = nullptr; }
    h->h_vn3.clear(); h->h_cn3.clear();
    h->use_triple = getenv("IBLDPC_NO_TRIPLE") == nullptr;
    {
        bool has3 = false;
        for (auto& c : h->vn_classes) has3 |= c.degree == 3;
        if (h->nib && h->use_triple && has3 && DV >= 3) {
            const int TT = T * T;
            std::vector<uint8_t> f3((size_t)imax * kTS * kTS * kTS, 0);
            for (int it = 0; it < imax; ++it) {
                const uint8_t* S0 = vn.data() + ((size_t)it * DV + 0) * TT;
                const uint8_t* S1 = vn.data() + ((size_t)it * DV + 1) * TT;
                const uint8_t* mrow = match ? mv.data() + ((size_t)it * DV + 2) * T : nullptr;
                uint8_t* dst = f3.data() + (size_t)it * kTS * kTS * kTS;
                for (int a = 0; a < T; ++a)
                    for (int b = 0; b < T; ++b)
                        for (int c = 0; c < T; ++c) {
                            int v = S1[S0[a * T + b] * T + c];
                            if (mrow) v = mrow[v];
                            dst[(b * kTS + c) * kTS + a] = (uint8_t)v;   // channel value minor (ib_triple_n4.cuh)
                        }
            }
            if ((rc = upload(&h->d_vn3, f3.data(), f3.size()))) return rc;
            h->h_vn3.swap(f3);
        }
    }
    // ... and of the first two check-node stages for the degree-6 class: F(x, y, z) = 4 * C_1( C_0(x, y), z ), index
    // (y*16 + z)*16 + x, for every table block (block 0 = iteration-0 tables)
    if (h->d_cn3) { CK(cudaFree(h->d_cn3)); h->d_cn3 = nullptr; }
    {
        bool has6 = false;   // a class the three-input-table kernels cover
        h->cn_tri_max_degree = getenv("IBLDPC_CN_TRI_MAX_DEGREE") ? atoi(getenv("IBLDPC_CN_TRI_MAX_DEGREE")) : 8;
        for (auto& c : h->cn_classes) has6 |= c.degree >= 6 && c.degree <= h->cn_tri_max_degree;
        if (h->nib && h->use_triple && h->use_pair && has6 && DC >= 6 && h->d_cn_pair != nullptr) {
            const int TT = T * T;
            std::vector<uint8_t> f3((size_t)imax * kTS * kTS * kTS, 0);
            for (int blk = 0; blk < imax; ++blk) {
                const uint8_t* C0 = cn.data() + ((size_t)blk * (DC - 2) + 0) * TT;
                const uint8_t* C1 = cn.data() + ((size_t)blk * (DC - 2) + 1) * TT;
                uint8_t* dst = f3.data() + (size_t)blk * kTS * kTS * kTS;
                for (int x = 0; x < T; ++x)
                    for (int y = 0; y < T; ++y)
                        for (int z = 0; z < T; ++z) dst[(y * kTS + z) * kTS + x] = (uint8_t)(4 * C1[C0[x * T + y] * T + z]);
            }
            if ((rc = upload(&h->d_cn3, f3.data(), f3.size()))) return rc;
            h->h_cn3.swap(f3);
        }
    }
    // Fused per-phase kernels (ib_phase_n4.cuh): default whenever the code's degree sets are instantiated; every switch
    // that selects a particular per-class kernel variant (A/B measurements, parity variants) keeps per-class launches.
    h->use_phase = getenv("IBLDPC_NO_PHASE") == nullptr && getenv("IBLDPC_PAIR_MIN_DEGREE") == nullptr &&
                   getenv("IBLDPC_VN_PAIR_MIN_DEGREE") == nullptr && h->vn_vec == 0 && h->cn_threads == 0 &&
                   h->vn_threads == 0 && h->vn_pair_threads == 0 && getenv("IBLDPC_NO_PLAN") == nullptr;
    if ((rc = phase_prepare(h))) return rc;
    // Batch-size policy between the whole-decode cooperative kernel, the fused per-phase kernels and one launch per degree
    // class (measured on B200, profiles/r02_small_and_mid_batches.txt).  Codes without fused kernels: cooperative kernel up
    // to 4096 frames.  Instantiated sets: one cooperative launch over the phase images (ib_coop_phase_kernel) up to 256
    // frames (lane = (node, word) mapping) and up to 2048 frames while the packed messages stay below 32 MB ((3,6) B=2048
    // 1.90 -> 1.71 ms, 802.11n 1.34 -> 1.15 ms; DVB-S2 n=64800 loses from 512 frames on: 5.25 -> 5.59 ms); above, the
    // 802.11n sets run the fused per-phase kernels at every batch size, the (3,6) and DVB-S2 sets the fused kernels up to
    // 4096 frames and per-class launches above.
    // IBLDPC_COOP_MAX_B / IBLDPC_PHASE_MID_MAX_B / IBLDPC_NO_COOP_PHASE override.
    if (const char* e = getenv("IBLDPC_COOP_MAX_B")) h->coop_max_frames = std::max(0LL, atoll(e));
    else h->coop_max_frames = h->phase == nullptr ? 4096 : 2048;
    h->coop_max_from_env = getenv("IBLDPC_COOP_MAX_B") != nullptr;
    h->no_coop_phase = getenv("IBLDPC_NO_COOP_PHASE") != nullptr;
    h->phase_mid_max_frames = getenv("IBLDPC_PHASE_MID_MAX_B") ? std::max(0LL, atoll(getenv("IBLDPC_PHASE_MID_MAX_B"))) : 4096;
    h->phase_off_midrange = getenv("IBLDPC_NO_PHASE") != nullptr;
    if ((rc = t32_prepare(h))) return rc;
    h->occ_cache.clear();
    h->have_luts = true;
    return IBLDPC_OK;
}

int ibldpc_decode_ib(ibldpc_handle h, const uint8_t* ch_dev, int64_t B, int imax, int early_term, uint8_t* out_dev,
                     int32_t* i_num_host, void* stream)
{
    int rc = check_decode_args(h, B, imax);
    if (rc) return rc;
    if (!ch_dev || !out_dev) return fail(IBLDPC_E_INVALID, "null buffer");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    cudaStream_t st = (cudaStream_t)stream;
    Workspace& w = h->ws[0];
    if (h->profiling) clear_events(h);
    const long long pitch = (B + 15) / 16 * 16;
    const bool aligned = (B % 16 == 0) && ((uintptr_t)ch_dev % 16 == 0) && ((uintptr_t)out_dev % 16 == 0);
    if (aligned) {
        rc = decode_ib_padded(h, w, ch_dev, pitch, B, imax, early_term, out_dev, st);
        if (rc) return rc;
    } else {
        if ((rc = ensure_pad(w, (size_t)h->N * pitch))) return rc;
        if ((rc = launch_pad<uint8_t>(ch_dev, w.padbuf_in, h->N, B, pitch, 0, st))) return rc;
        if ((rc = decode_ib_padded(h, w, w.padbuf_in, pitch, B, imax, early_term, w.padbuf_out, st))) return rc;
        if ((rc = launch_unpad<uint8_t>(w.padbuf_out, out_dev, h->N, B, pitch, st))) return rc;
    }
    if (i_num_host) return read_back_status(w, st, i_num_host);
    return IBLDPC_OK;
}

int ibldpc_decode_ib_perframe(ibldpc_handle h, const uint8_t* ch_dev, int64_t B, int imax, uint8_t* out_dev,
                              int32_t* i_num_frames_dev, void* stream)
{
    int rc = check_decode_args(h, B, imax);
    if (rc) return rc;
    if (!ch_dev || !out_dev) return fail(IBLDPC_E_INVALID, "null buffer");
    if (!(h->fast && h->nib) || h->phase == nullptr)
        return fail(IBLDPC_E_STATE, "per-frame early termination needs the packed-nibble family and the fused per-phase kernels "
                                    "(even |T| <= 16, instantiated degree sets: 802.11n, DVB-S2 rate 1/2, regular (3,6))");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    cudaStream_t st = (cudaStream_t)stream;
    Workspace& w = h->ws[0];
    if (h->profiling) clear_events(h);
    const long long pitch = (B + 15) / 16 * 16;
    const bool aligned = (B % 16 == 0) && ((uintptr_t)ch_dev % 16 == 0) && ((uintptr_t)out_dev % 16 == 0);
    h->pf_request = true;
    h->pf_inum = i_num_frames_dev;
    if (aligned) {
        rc = decode_ib_padded(h, w, ch_dev, pitch, B, imax, 1, out_dev, st);
    } else {
        rc = ensure_pad(w, (size_t)h->N * pitch);
        if (!rc) rc = launch_pad<uint8_t>(ch_dev, w.padbuf_in, h->N, B, pitch, 0, st);
        if (!rc) rc = decode_ib_padded(h, w, w.padbuf_in, pitch, B, imax, 1, w.padbuf_out, st);
        if (!rc) rc = launch_unpad<uint8_t>(w.padbuf_out, out_dev, h->N, B, pitch, st);
    }
    h->pf_request = false;
    h->pf_inum = nullptr;
    return rc;
}

int ibldpc_last_i_num(ibldpc_handle h, int32_t* i_num_host)
{
    if (!h || !i_num_host) return fail(IBLDPC_E_INVALID, "null argument");
    Workspace& w = h->ws[h->last_ws];
    if (!w.inum) return fail(IBLDPC_E_STATE, "no decode has been issued on this handle");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    return read_back_status(w, h->last_stream, i_num_host);
}

int ibldpc_decode_ib_host(ibldpc_handle h, const uint8_t* ch_host, int64_t B, int imax, int early_term,
                          uint8_t* out_host, int32_t* i_num_host)
{
    int rc = check_decode_args(h, B, imax);
    if (rc) return rc;
    if (!ch_host || !out_host) return fail(IBLDPC_E_INVALID, "null buffer");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    // Early termination is a property of the whole call (all B frames), so it cannot be chunked.
    const std::vector<int64_t> widths = host_chunk_schedule(B, h->N, h->host_chunk, early_term != 0);
    int64_t chunk = 0;
    for (int64_t wdt : widths) chunk = std::max(chunk, wdt);
    const long long cpitch = (chunk + 15) / 16 * 16;
    const int nslots = widths.size() > 1 ? 2 : 1;
    for (int s = 0; s < nslots; ++s) {
        Workspace& w = h->ws[s];
        if (!w.stream) CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
        const size_t need = (size_t)h->N * cpitch;
        if (w.stage_bytes < need) {
            if (w.stage_in) CK(cudaFree(w.stage_in));
            if (w.stage_out) CK(cudaFree(w.stage_out));
            w.stage_in = w.stage_out = nullptr;
            w.stage_bytes = 0;
            if (cudaMalloc((void**)&w.stage_in, need) != cudaSuccess || cudaMalloc((void**)&w.stage_out, need) != cudaSuccess)
                return fail(IBLDPC_E_NOMEM, "cudaMalloc of staging buffers failed");
            CK(cudaMemset(w.stage_in, 0, need));
            w.stage_bytes = need;
        }
    }
    if (h->profiling) clear_events(h);
    int slot = 0;
    int64_t off = 0;
    for (int64_t wd : widths) {
        Workspace& w = h->ws[slot];
        const long long pitch = (wd + 15) / 16 * 16;
        CK(cudaMemcpy2DAsync(w.stage_in, (size_t)pitch, ch_host + off, (size_t)B, (size_t)wd, (size_t)h->N,
                             cudaMemcpyHostToDevice, w.stream));
        if ((rc = decode_ib_padded(h, w, w.stage_in, pitch, wd, imax, early_term, w.stage_out, w.stream))) return rc;
        CK(cudaMemcpy2DAsync(out_host + off, (size_t)B, w.stage_out, (size_t)pitch, (size_t)wd, (size_t)h->N,
                             cudaMemcpyDeviceToHost, w.stream));
        off += wd;
        slot ^= (nslots - 1);
    }
    int first_rc = IBLDPC_OK;   // read (and thereby clear) the status of every slot, report the first failure
    for (int s = nslots - 1; s >= 0; --s)
        if ((rc = read_back_status(h->ws[s], h->ws[s].stream, s == 0 ? i_num_host : nullptr)) && !first_rc) first_rc = rc;
    return first_rc;
}

// ---- packed host buffers: nibble-packed channel values in, bit-packed hard decisions of the first `rows` rows out
int ibldpc_decode_ib_host_packed(ibldpc_handle h, const uint8_t* ch4_host, int64_t B, int imax, int early_term,
                                 uint8_t* bits_host, int64_t rows, int32_t* i_num_host)
{
    int rc = check_decode_args(h, B, imax);
    if (rc) return rc;
    if (!ch4_host || !bits_host) return fail(IBLDPC_E_INVALID, "null buffer");
    if (rows < 0 || rows > h->N) return fail(IBLDPC_E_INVALID, "rows must lie in [0, n_var]");
    if (h->T > 16 || h->Tc > 16) return fail(IBLDPC_E_INVALID, "packed host buffers need cardinalities <= 16 (one nibble per cluster index)");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    const std::vector<int64_t> widths = host_chunk_schedule(B, h->N, h->host_chunk, early_term != 0);
    int64_t chunk = 0;
    for (int64_t wdt : widths) chunk = std::max(chunk, wdt);
    const long long cpitch = (chunk + 15) / 16 * 16;
    const int nslots = widths.size() > 1 ? 2 : 1;
    const bool nib = h->fast && h->nib;
    const size_t src_pitch = (size_t)((B + 1) / 2), dst_pitch = (size_t)((B + 7) / 8);
    for (int s = 0; s < nslots; ++s) {
        Workspace& w = h->ws[s];
        if (!w.stream) CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
        if ((rc = ensure_ws_common(h, w))) return rc;
        if ((rc = ensure_n4_buffers(h, w, chunk))) return rc;
        const size_t need = (size_t)h->N * cpitch;
        if (w.stage_bytes < need) {
            if (w.stage_in) CK(cudaFree(w.stage_in));
            if (w.stage_out) CK(cudaFree(w.stage_out));
            w.stage_in = w.stage_out = nullptr;
            w.stage_bytes = 0;
            if (cudaMalloc((void**)&w.stage_in, need) != cudaSuccess || cudaMalloc((void**)&w.stage_out, need) != cudaSuccess)
                return fail(IBLDPC_E_NOMEM, "cudaMalloc of staging buffers failed");
            CK(cudaMemset(w.stage_in, 0, need));
            w.stage_bytes = need;
        }
        void* p = w.stage_bits;
        rc = ensure(&p, &w.stage_bits_bytes, (size_t)std::max<int64_t>(rows, 1) * (size_t)((chunk + 31) / 32) * 4);
        w.stage_bits = (uint8_t*)p;
        if (rc) return rc;
    }
    if (h->profiling) clear_events(h);
    int slot = 0;
    int64_t off = 0;
    for (int64_t wd : widths) {
        Workspace& w = h->ws[slot];
        if (off % 8) return fail(IBLDPC_E_INVALID, "host chunk size must be a multiple of 8 frames");
        const long long pitch = (wd + 15) / 16 * 16;
        const long long pitch4 = ((wd + 1) / 2 + 15) / 16 * 16;
        CK(cudaMemcpy2DAsync(w.ch4, (size_t)pitch4, ch4_host + off / 2, src_pitch, (size_t)((wd + 1) / 2), (size_t)h->N,
                             cudaMemcpyHostToDevice, w.stream));
        if (nib) {
            if ((rc = decode_ib_padded(h, w, nullptr, pitch, wd, imax, early_term, w.stage_out, w.stream, true))) return rc;
        } else {
            const long long n = (long long)h->N * pitch;
            unpack_n4_kernel<<<(int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 16), 256, 0, w.stream>>>(
                w.ch4, w.stage_in, h->N, wd, pitch, (uint32_t)pitch4);
            if ((rc = decode_ib_padded(h, w, w.stage_in, pitch, wd, imax, early_term, w.stage_out, w.stream))) return rc;
        }
        if (rows > 0) {
            const long long wpr = (wd + 31) / 32;
            const long long n = rows * wpr;
            harddecision_bits_kernel<<<(int)std::min<long long>((n + 255) / 256, (long long)h->sm_count * 16), 256, 0, w.stream>>>(
                w.stage_out, (int)rows, wd, pitch, h->T / 2, reinterpret_cast<uint32_t*>(w.stage_bits), wpr);
            h->last_launches++;
            CK(cudaMemcpy2DAsync(bits_host + off / 8, dst_pitch, w.stage_bits, (size_t)wpr * 4, (size_t)((wd + 7) / 8), (size_t)rows,
                                 cudaMemcpyDeviceToHost, w.stream));
        }
        off += wd;
        slot ^= (nslots - 1);
    }
    CK(cudaGetLastError());
    int first_rc = IBLDPC_OK;   // read (and thereby clear) the status of every slot, report the first failure
    for (int s = nslots - 1; s >= 0; --s)
        if ((rc = read_back_status(h->ws[s], h->ws[s].stream, s == 0 ? i_num_host : nullptr)) && !first_rc) first_rc = rc;
    return first_rc;
}

// ---- the reference's own host contract: int32 cluster indices in, int32 cluster indices out
// (discrete_LDPC_decoder.py:207-209 uploads received_blocks.astype(np.int32), :292-295 returns the int32 output).
// The narrowing to uint8 / widening to int32 runs on host threads, chunk by chunk, into pinned staging buffers,
// overlapped with the copies and the decode of the neighbouring chunk.

int ibldpc_decode_ib_host_i32(ibldpc_handle h, const int32_t* ch_host, int64_t B, int imax, int early_term,
                              int32_t* out_host, int32_t* i_num_host)
{
    int rc = check_decode_args(h, B, imax);
    if (rc) return rc;
    if (!ch_host || !out_host) return fail(IBLDPC_E_INVALID, "null buffer");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    const std::vector<int64_t> widths = host_chunk_schedule(B, h->N, h->host_chunk, early_term != 0);
    int64_t chunk = 0;
    for (int64_t wdt : widths) chunk = std::max(chunk, wdt);
    const long long cpitch = (chunk + 15) / 16 * 16;
    const int nslots = widths.size() > 1 ? 2 : 1;
    const size_t need = (size_t)h->N * cpitch;
    for (int s = 0; s < nslots; ++s) {
        Workspace& w = h->ws[s];
        if (!w.stream) CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
        if (!w.done_ev) CK(cudaEventCreateWithFlags(&w.done_ev, cudaEventDisableTiming));
        if (w.stage_bytes < need) {
            if (w.stage_in) CK(cudaFree(w.stage_in));
            if (w.stage_out) CK(cudaFree(w.stage_out));
            w.stage_in = w.stage_out = nullptr;
            w.stage_bytes = 0;
            if (cudaMalloc((void**)&w.stage_in, need) != cudaSuccess || cudaMalloc((void**)&w.stage_out, need) != cudaSuccess)
                return fail(IBLDPC_E_NOMEM, "cudaMalloc of staging buffers failed");
            CK(cudaMemset(w.stage_in, 0, need));
            w.stage_bytes = need;
        }
        if (w.pin_bytes < need) {
            if (w.pin_in) CK(cudaFreeHost(w.pin_in));
            if (w.pin_out) CK(cudaFreeHost(w.pin_out));
            w.pin_in = w.pin_out = nullptr;
            w.pin_bytes = 0;
            if (cudaHostAlloc((void**)&w.pin_in, need, cudaHostAllocDefault) != cudaSuccess ||
                cudaHostAlloc((void**)&w.pin_out, need, cudaHostAllocDefault) != cudaSuccess)
                return fail(IBLDPC_E_NOMEM, "cudaHostAlloc of pinned staging buffers failed");
            w.pin_bytes = need;
        }
    }
    if (h->profiling) clear_events(h);
    static const int nthreads = [] {
        if (const char* e = getenv("IBLDPC_HOST_THREADS")) return std::max(1, atoi(e));
        return (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    }();
    const int N = h->N, Tc = h->Tc;
    struct Pending { bool busy = false; int64_t off = 0, wd = 0; long long pitch = 0; };
    Pending pend[2];
    auto drain = [&](int s) -> int {
        Pending& q = pend[s];
        if (!q.busy) return IBLDPC_OK;
        Workspace& w = h->ws[s];
        CK(cudaEventSynchronize(w.done_ev));
        const uint8_t* src = w.pin_out;
        const int64_t off = q.off, wd = q.wd;
        const long long pitch = q.pitch;
        parallel_rows(N, nthreads, [=](int64_t lo, int64_t hi) {
            for (int64_t r = lo; r < hi; ++r) {
                const uint8_t* a = src + r * pitch;
                int32_t* o = out_host + r * B + off;
                for (int64_t f = 0; f < wd; ++f) o[f] = a[f];
            }
        });
        q.busy = false;
        return IBLDPC_OK;
    };
    int slot = 0;
    int64_t off = 0;
    for (int64_t wd : widths) {
        Workspace& w = h->ws[slot];
        if ((rc = drain(slot))) return rc;
        const long long pitch = (wd + 15) / 16 * 16;
        std::vector<int> bad((size_t)nthreads + 1, 0);
        uint8_t* dst = w.pin_in;
        int* badp = bad.data();
        const int64_t per = (N + nthreads - 1) / nthreads;
        parallel_rows(N, nthreads, [=](int64_t lo, int64_t hi) {
            int any = 0;
            for (int64_t r = lo; r < hi; ++r) {
                const int32_t* a = ch_host + r * B + off;
                uint8_t* o = dst + r * pitch;
                for (int64_t f = 0; f < wd; ++f) {
                    const int32_t v = a[f];
                    any |= (v < 0) | (v >= Tc);
                    o[f] = (uint8_t)v;
                }
                for (int64_t f = wd; f < pitch; ++f) o[f] = 0;
            }
            badp[lo / per] = any;
        });
        for (int b : bad)
            if (b) {
                for (int s = 0; s < nslots; ++s) cudaStreamSynchronize(h->ws[s].stream);
                return fail(IBLDPC_E_INVALID, "channel cluster indices must lie in [0, cardinality_T_channel)");
            }
        CK(cudaMemcpyAsync(w.stage_in, w.pin_in, (size_t)N * pitch, cudaMemcpyHostToDevice, w.stream));
        if ((rc = decode_ib_padded(h, w, w.stage_in, pitch, wd, imax, early_term, w.stage_out, w.stream))) return rc;
        CK(cudaMemcpyAsync(w.pin_out, w.stage_out, (size_t)N * pitch, cudaMemcpyDeviceToHost, w.stream));
        CK(cudaEventRecord(w.done_ev, w.stream));
        pend[slot].busy = true; pend[slot].off = off; pend[slot].wd = wd; pend[slot].pitch = pitch;
        off += wd;
        slot ^= (nslots - 1);
    }
    // drain in issue order: `slot` now names the older of the two pending chunks
    if ((rc = drain(slot))) return rc;
    if (nslots > 1 && (rc = drain(slot ^ 1))) return rc;
    int first_rc = IBLDPC_OK;   // read (and thereby clear) the status of every slot, report the first failure
    for (int s = nslots - 1; s >= 0; --s)
        if ((rc = read_back_status(h->ws[s], h->ws[s].stream, s == 0 ? i_num_host : nullptr)) && !first_rc) first_rc = rc;
    return first_rc;
}

int ibldpc_decode_llr(ibldpc_handle h, int algo, int dtype, const void* ch_dev, int64_t B, int imax, int early_term,
                      void* out_dev, int32_t* i_num_host, void* stream)
{
    if (!h) return fail(IBLDPC_E_INVALID, "null handle");
    if (!ch_dev || !out_dev) return fail(IBLDPC_E_INVALID, "null buffer");
    if (B <= 0 || B > 0x7fffffffLL - 1024) return fail(IBLDPC_E_INVALID, "bad B");
    if (imax < 1 || imax >= kMaxIter) return fail(IBLDPC_E_INVALID, "imax out of range");
    if (algo != IBLDPC_ALGO_MINSUM && algo != IBLDPC_ALGO_BP) return fail(IBLDPC_E_INVALID, "unknown algorithm");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == IBLDPC_F32)
        return decode_llr_typed<float>(h, algo, (const float*)ch_dev, B, imax, early_term, (float*)out_dev, i_num_host, st);
    if (dtype == IBLDPC_F64)
        return decode_llr_typed<double>(h, algo, (const double*)ch_dev, B, imax, early_term, (double*)out_dev, i_num_host, st);
    return fail(IBLDPC_E_INVALID, "dtype must be IBLDPC_F32 or IBLDPC_F64");
}

int ibldpc_decode_llr_layered(ibldpc_handle h, int algo, int dtype, const void* ch_dev, int64_t B, int imax, int early_term,
                              void* out_dev, int32_t* i_num_host, void* stream)
{
    if (!h) return fail(IBLDPC_E_INVALID, "null handle");
    if (!ch_dev || !out_dev) return fail(IBLDPC_E_INVALID, "null buffer");
    if (B <= 0 || B > 0x7fffffffLL - 1024) return fail(IBLDPC_E_INVALID, "bad B");
    if (imax < 1 || imax >= kMaxIter) return fail(IBLDPC_E_INVALID, "imax out of range");
    if (algo != IBLDPC_ALGO_MINSUM && algo != IBLDPC_ALGO_BP) return fail(IBLDPC_E_INVALID, "unknown algorithm");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == IBLDPC_F32)
        return decode_llr_layered_typed<float>(h, algo, (const float*)ch_dev, B, imax, early_term, (float*)out_dev, i_num_host, st);
    if (dtype == IBLDPC_F64)
        return decode_llr_layered_typed<double>(h, algo, (const double*)ch_dev, B, imax, early_term, (double*)out_dev, i_num_host, st);
    return fail(IBLDPC_E_INVALID, "dtype must be IBLDPC_F32 or IBLDPC_F64");
}

int ibldpc_layer_count(ibldpc_handle h, int32_t* n_layers)
{
    if (!h || !n_layers) return fail(IBLDPC_E_INVALID, "null argument");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    int rc = layered_prepare(h);
    if (rc) return rc;
    *n_layers = layered_count(h);
    return IBLDPC_OK;
}

int ibldpc_count_errors_u8(int device, const uint8_t* out_dev, int64_t rows, int64_t B, int threshold,
                           const uint8_t* ref_bits_dev, int64_t* counters_host, void* stream)
{
    if (threshold < 0 || threshold > 255) return fail(IBLDPC_E_INVALID, "threshold out of range");
    return count_errors_common<uint8_t>(device, out_dev, rows, B, (uint8_t)threshold, ref_bits_dev, counters_host, nullptr, stream);
}

int ibldpc_count_errors_u8_async(int device, const uint8_t* out_dev, int64_t rows, int64_t B, int threshold,
                                 const uint8_t* ref_bits_dev, int64_t* counters_dev, void* stream)
{
    if (threshold < 0 || threshold > 255) return fail(IBLDPC_E_INVALID, "threshold out of range");
    if (!counters_dev) return fail(IBLDPC_E_INVALID, "null device counters");
    return count_errors_common<uint8_t>(device, out_dev, rows, B, (uint8_t)threshold, ref_bits_dev, nullptr, counters_dev, stream);
}

int ibldpc_count_errors_llr(int device, const void* out_dev, int dtype, int64_t rows, int64_t B,
                            const uint8_t* ref_bits_dev, int64_t* counters_host, void* stream)
{
    if (dtype == IBLDPC_F32)
        return count_errors_common<float>(device, (const float*)out_dev, rows, B, 0.f, ref_bits_dev, counters_host, nullptr, stream);
    if (dtype == IBLDPC_F64)
        return count_errors_common<double>(device, (const double*)out_dev, rows, B, 0.0, ref_bits_dev, counters_host, nullptr, stream);
    return fail(IBLDPC_E_INVALID, "dtype must be IBLDPC_F32 or IBLDPC_F64");
}

int ibldpc_count_errors_llr_async(int device, const void* out_dev, int dtype, int64_t rows, int64_t B,
                                  const uint8_t* ref_bits_dev, int64_t* counters_dev, void* stream)
{
    if (!counters_dev) return fail(IBLDPC_E_INVALID, "null device counters");
    if (dtype == IBLDPC_F32)
        return count_errors_common<float>(device, (const float*)out_dev, rows, B, 0.f, ref_bits_dev, nullptr, counters_dev, stream);
    if (dtype == IBLDPC_F64)
        return count_errors_common<double>(device, (const double*)out_dev, rows, B, 0.0, ref_bits_dev, nullptr, counters_dev, stream);
    return fail(IBLDPC_E_INVALID, "dtype must be IBLDPC_F32 or IBLDPC_F64");
}

int ibldpc_quantize(int device, const double* x_dev, int64_t n, const double* limits_host, int card, uint8_t* out_dev,
                    void* stream)
{
    if (card > 256) return fail(IBLDPC_E_INVALID, "card must be <= 256 for uint8 clusters");
    return quantize_common<0>(device, x_dev, n, limits_host, card, nullptr, 0, 0, 0, out_dev, stream);
}

int ibldpc_quantize_llr(int device, const double* x_dev, int64_t n, const double* limits_host, int card,
                        const double* llr_host, int dtype, void* out_dev, void* stream)
{
    if (dtype != IBLDPC_F32 && dtype != IBLDPC_F64) return fail(IBLDPC_E_INVALID, "bad dtype");
    return quantize_common<0>(device, x_dev, n, limits_host, card, llr_host, dtype == IBLDPC_F32 ? 1 : 2, 0, 0, out_dev, stream);
}

int ibldpc_sample_direct(int device, const double* cdf_host, int card, uint64_t seed, uint64_t offset, int64_t n,
                         uint8_t* out_dev, void* stream)
{
    if (card > 257) return fail(IBLDPC_E_INVALID, "card must be <= 257");
    return quantize_common<1>(device, nullptr, n, cdf_host, card, nullptr, 0, seed, offset, out_dev, stream);
}

int ibldpc_sample_direct_llr(int device, const double* cdf_host, int card, const double* llr_host, uint64_t seed,
                             uint64_t offset, int64_t n, int dtype, void* out_dev, void* stream)
{
    if (dtype != IBLDPC_F32 && dtype != IBLDPC_F64) return fail(IBLDPC_E_INVALID, "bad dtype");
    return quantize_common<1>(device, nullptr, n, cdf_host, card, llr_host, dtype == IBLDPC_F32 ? 1 : 2, seed, offset, out_dev, stream);
}

int ibldpc_uniform(int device, uint64_t seed, uint64_t offset, int64_t n, double* out_dev, void* stream)
{
    return quantize_common<1>(device, nullptr, n, nullptr, 1, nullptr, 3, seed, offset, out_dev, stream);
}

int ibldpc_info(ibldpc_handle h, int32_t* which4)
{
    if (!h || !which4) return fail(IBLDPC_E_INVALID, "null argument");
    which4[0] = h->fast ? (h->nib ? 2 : 1) : h->t32 ? 3 : 0;   // 0 generic, 1 uint8, 2 packed-nibble, 3 |T| <= 32 family
    which4[1] = h->last_launches;
    which4[2] = h->last_grid;
    which4[3] = h->last_smem;
    return IBLDPC_OK;
}

int ibldpc_set_profiling(ibldpc_handle h, int on)
{
    if (!h) return fail(IBLDPC_E_INVALID, "null handle");
    h->profiling = on != 0;
    if (!on) clear_events(h);
    return IBLDPC_OK;
}

int ibldpc_set_host_chunk(ibldpc_handle h, int frames)
{
    if (!h || (frames != 0 && frames < 16)) return fail(IBLDPC_E_INVALID, "bad chunk");
    h->host_chunk = frames;
    return IBLDPC_OK;
}

int ibldpc_phase_times(ibldpc_handle h, float* ms3, int32_t* launches3)
{
    if (!h || !ms3 || !launches3) return fail(IBLDPC_E_INVALID, "null argument");
    DeviceGuard guard_(h->device);
    CK(guard_.err);
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < 3; ++i) { ms3[i] = 0.f; launches3[i] = 0; }
    for (auto& e : h->events) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e.a, e.b));
        ms3[e.phase] += ms;
        launches3[e.phase] += 1;
    }
    return IBLDPC_OK;
}

int ibldpc_plan_geometry(int64_t resident_ctas, int warps_per_cta, int tiles, int n_nodes, int32_t* out3)
{
    if (!out3 || resident_ctas < 1 || warps_per_cta < 8 || warps_per_cta % 8 || tiles < 1 || n_nodes < 1)
        return fail(IBLDPC_E_INVALID, "bad arguments");
    int tpc, tg, gx;
    plan_geometry(resident_ctas, warps_per_cta, tiles, n_nodes, false, &tpc, &tg, &gx);
    out3[0] = tpc; out3[1] = tg; out3[2] = gx;
    return IBLDPC_OK;
}

int ibldpc_host_chunk_schedule(int64_t B, int64_t n_var, int64_t host_chunk, int early_term, int64_t* widths, int capacity)
{
    if (B < 1 || n_var < 1 || host_chunk < 0 || capacity < 0 || (capacity > 0 && !widths))
        return fail(IBLDPC_E_INVALID, "bad arguments");
    const std::vector<int64_t> w = host_chunk_schedule(B, n_var, host_chunk, early_term != 0);
    for (size_t i = 0; i < w.size() && (int)i < capacity; ++i) widths[i] = w[i];
    return (int)w.size();
}

int ibldpc_destroy(ibldpc_handle h)
{
    if (!h) return IBLDPC_OK;
    DeviceGuard guard_(h->device);
    cudaDeviceSynchronize();
    ibldpc_nccl_finalize(h);
    phase_free(h);
    t32_free(h);
    layered_free(h);
    clear_events(h);
    for (int* p : {h->d_sc, h->d_dc, h->d_tc, h->d_sv, h->d_dv, h->d_tv, h->d_vidx})
        if (p) cudaFree(p);
    for (uint8_t* p : {h->d_cn8, h->d_vn8, h->d_mc8, h->d_mv8, h->d_cn_pair, h->d_vn_pair, h->d_vn3, h->d_cn3})
        if (p) cudaFree(p);
    for (auto& c : h->cn_classes) if (c.d_nodes) cudaFree(c.d_nodes);
    for (auto& c : h->vn_classes) if (c.d_nodes) cudaFree(c.d_nodes);
    for (auto& w : h->ws) {
        for (void* p : {(void*)w.msg, (void*)w.ch4, w.cin, w.vin, (void*)w.padbuf_in, (void*)w.padbuf_out, (void*)w.stage_in,
                        (void*)w.stage_out, (void*)w.flags, (void*)w.inum})
            if (p) cudaFree(p);
        for (void* p : {(void*)w.stage_bits, (void*)w.pf_idx})
            if (p) cudaFree(p);
        if (w.pin_in) cudaFreeHost(w.pin_in);
        if (w.pin_out) cudaFreeHost(w.pin_out);
        if (w.done_ev) cudaEventDestroy(w.done_ev);
        if (w.stream) cudaStreamDestroy(w.stream);
        if (w.fork_ev) cudaEventDestroy(w.fork_ev);
        for (int i = 0; i < 7; ++i) {
            if (w.aux[i]) cudaStreamDestroy(w.aux[i]);
            if (w.join_ev[i]) cudaEventDestroy(w.join_ev[i]);
        }
    }
    delete h;
    return IBLDPC_OK;
}

}  // extern "C"
