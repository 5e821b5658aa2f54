// ib_perframe.cu -- opt-in PER-FRAME early termination with frame compaction.
//
// The reference stops a whole batch at once (discrete_LDPC_decoder.py:233,273: `while i_num < imax and not
// syndrome_zero` on the SUM of the syndrome over all msg_at_time frames).  Decoding every frame on its own
// (msg_at_time = 1) stops each frame as soon as ITS syndrome is zero; that is the result this mode reproduces for a
// whole batch in one call -- outputs and per-frame i_num equal to the reference run with one frame per call -- while
// only the frames that still iterate cost anything:
//   * every check-node phase records, per frame, whether a check failed (bit 4f of pf_fsyn[word]);
//   * pf_update_kernel retires the alive frames without a failed check: it stores their i_num, clears them from the
//     `alive` nibble mask and hands every one of them a dense RESULT SLOT (group g = pass g - 1 owns the slots
//     [gstart[g], gstart[g + 1]));
//   * from then on both phases leave the messages of such a frame untouched (cn_node_n4 / PhaseItem merge under the
//     alive mask): the frame stays FROZEN in the state calc_varnode_output would read
//     (discrete_LDPC_decoder.py:280-287);
//   * ib_phase_pfdecide_kernel decides the frozen frames group by group with the tables of their own iteration
//     (i_num - 1), gathering the eight frames of a result word from their columns: every frame is decided exactly once,
//     densely, however the convergence passes interleave inside a word;
//   * after the pending groups are decided, the batch is compacted IN PLACE when that pays: every dead column of the new
//     front [0, n_alive) is filled with an alive column of the tail (only the holes move: one CTA per row stages the row
//     in shared memory, patches the hole nibbles and writes the front back).  It pays when the passes' worth of work
//     spent on dead columns since the last compaction has reached the cost of one, and the rest of the schedule is long
//     enough to win it back (ski-rental rule, evaluated on the device);
//   * one kernel at the end permutes the result nibbles from slot order back to the caller's frame order (uint8).
// Nothing synchronises with the host: all launches of the schedule are issued up front and read what is left to do
// from a device-side PfState.  Runs on the fused per-phase kernels (ib_phase_n4.cuh), i.e. for the degree sets those
// are instantiated for.
#include <algorithm>
#include <cstdlib>
#include <string>
#include <vector>

#include "ibldpc_internal.h"
#include "ib_phase_sets.h"

namespace ibldpc {

// defined in ib_phase.cu
struct PhaseImages;
const PhaseSetOps* phase_ops_of(const ibldpc_decoder* h);
void phase_fill_args(const ibldpc_decoder* h, int mode, int index, PhaseArgs& q, size_t* smem);
int phase_set_attributes(ibldpc_decoder* h);

namespace {

constexpr int kGatherThreads = 512;
constexpr int kExpandThreads = 512;
constexpr size_t kRowSmemMax = 200 * 1024;   // a row of active columns must fit in shared memory to be compacted in place

__global__ void pf_init_kernel(PfState* st, int B, int pitch4, uint32_t* alive, uint32_t* fsyn, int* idx0, int* gstart, int words)
{
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        const int nv = B - 8 * w;
        alive[w] = nv >= 8 ? 0xffffffffu : nv <= 0 ? 0u : ((1u << (4 * nv)) - 1u);
        fsyn[w] = 0u;
    }
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < words * 8; f += gridDim.x * blockDim.x) idx0[f] = f < B ? f : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_act = B; st->act_pitch = pitch4; st->n_holes = 0; st->n_alive = B; st->done = 0; st->do_compact = 0;
        st->new_n = B; st->alive_acc = 0; st->blocks_done = 0u; st->blocks_done2 = 0u; st->fin_count = 0; st->waste = 0.f;
        gstart[0] = 0;
    }
}

// After the check-node phase of pass `it`: the alive frames without a failed check (all alive frames in the last pass)
// are retired -- i_num = it + 2 (passes executed + 1), a result slot of group it + 1 each, cleared from `alive` -- and the
// rest is counted.  The last block closes the group (slot count rounded up to a result word) and raises `done`.
__global__ void pf_update_kernel(PfState* st, uint32_t* alive, uint32_t* fsyn, const int* idx, int* fin, int* slot_of, int* gstart,
                                 int32_t* inum_frames, int* inum_batch, int it, int last, int imax)
{
    if (st->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) *inum_batch = it + 2;   // i_num of the last frame to finish = the batch's i_num
    const int n_act = st->n_act;
    const int words = (n_act + 7) >> 3;
    int cnt = 0;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        const uint32_t al = alive[w];
        const uint32_t failed = (fsyn[w] & 0x11111111u) * 15u;
        const uint32_t cv = last ? al : (al & ~failed);
        fsyn[w] = 0u;
        cnt += __popc(al & ~cv) >> 2;
        if (cv != 0u) {
            alive[w] = al & ~cv;
            int slot = atomicAdd(&st->fin_count, __popc(cv) >> 2);
#pragma unroll
            for (int f = 0; f < 8; ++f)
                if ((cv >> (4 * f)) & 1u) {
                    const int orig = idx[8 * w + f];
                    fin[slot] = 8 * w + f;
                    slot_of[orig] = slot;
                    if (inum_frames != nullptr) inum_frames[orig] = it + 2;
                    ++slot;
                }
        }
    }
    // block reduction, last block publishes
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt) atomicAdd(&st->alive_acc, s_cnt);
        __threadfence();
        const unsigned done = atomicAdd(&st->blocks_done, 1u);
        if (done == gridDim.x - 1) {
            __threadfence();
            const int n = atomicExch(&st->alive_acc, 0);
            st->n_alive = n;
            st->blocks_done = 0u;
            int fc = atomicAdd(&st->fin_count, 0);
            while (fc & 7) fin[fc++] = -1;            // padding slots of the group's last result word
            st->fin_count = fc;
            gstart[it + 2] = fc;                      // group it + 1 = [gstart[it + 1], gstart[it + 2])
            st->waste += (float)(n_act - n) / (float)n_act;
            if (n == 0) {                             // nothing left: the groups of the remaining passes are empty
                for (int g = it + 3; g <= imax; ++g) gstart[g] = fc;
                st->done = 1;
            }
        }
    }
}

// One CTA: decide whether a compaction pays now (see the header) and, if so, pair every dead column of the new front
// [0, n_alive) -- a HOLE -- with an alive column of the tail [n_alive, n_act): hole_dst[i] receives column hole_src[i].
// Only the holes move; every other alive column stays where it is.
__global__ void __launch_bounds__(1024) pf_scan_kernel(PfState* st, const uint32_t* alive, int* hole_dst, int* hole_src,
                                                       int remaining, float cost, int max_act_pitch)
{
    __shared__ int s_f[1024], s_t[1024];
    if (st->done) return;
    const int n_act = st->n_act, n_alive = st->n_alive;
    const float dead = (float)(n_act - n_alive) / (float)n_act;
    const bool pays = n_alive > 0 && n_act > 1024 && st->act_pitch <= max_act_pitch && (st->waste >= cost || dead >= 0.25f) &&
                      dead * (float)remaining >= cost;
    if (!pays) {
        if (threadIdx.x == 0) st->do_compact = 0;
        return;
    }
    const int words = (n_act + 7) >> 3;
    const int per = (words + 1023) / 1024;
    const int w0 = threadIdx.x * per, w1 = min(words, w0 + per);
    int cf = 0, ct = 0;
    for (int w = w0; w < w1; ++w) {
        const uint32_t al = alive[w];
        for (int f = 0; f < 8; ++f) {
            const int c = 8 * w + f;
            const bool a1 = (al >> (4 * f)) & 1u;
            cf += (c < n_alive && !a1) ? 1 : 0;
            ct += (c >= n_alive && a1) ? 1 : 0;
        }
    }
    s_f[threadIdx.x] = cf;
    s_t[threadIdx.x] = ct;
    __syncthreads();
    // inclusive scans (Hillis-Steele, 1024 entries)
    for (int off = 1; off < 1024; off <<= 1) {
        const int vf = threadIdx.x >= off ? s_f[threadIdx.x - off] : 0;
        const int vt = threadIdx.x >= off ? s_t[threadIdx.x - off] : 0;
        __syncthreads();
        s_f[threadIdx.x] += vf;
        s_t[threadIdx.x] += vt;
        __syncthreads();
    }
    int pf = s_f[threadIdx.x] - cf, pt = s_t[threadIdx.x] - ct;
    for (int w = w0; w < w1; ++w) {
        const uint32_t al = alive[w];
        for (int f = 0; f < 8; ++f) {
            const int c = 8 * w + f;
            const bool a1 = (al >> (4 * f)) & 1u;
            if (c < n_alive && !a1) hole_dst[pf++] = c;
            if (c >= n_alive && a1) hole_src[pt++] = c;
        }
    }
    if (threadIdx.x == 1023) {
        st->n_holes = s_f[1023];          // == s_t[1023]: dead columns in front = alive columns behind it
        st->new_n = n_alive;
        st->do_compact = 1;
    }
}

// In-place compaction: row by row (message rows, then channel rows), the active prefix of a row is staged in shared
// memory, the holes of the new front are patched with the nibbles of the tail columns, and the front is written back.
__global__ void __launch_bounds__(kGatherThreads) pf_gather_kernel(const PfState* st, uint8_t* msg, uint8_t* ch, int n_edge, int n_var,
                                                                  uint32_t pitch, const int* __restrict__ hole_dst,
                                                                  const int* __restrict__ hole_src)
{
    extern __shared__ __align__(16) uint8_t s_row[];
    if (st->done || !st->do_compact) return;
    const int new_n = st->new_n, n_holes = st->n_holes;
    const int act16 = st->act_pitch >> 4;
    const int new16 = (((new_n + 1) / 2 + 15) / 16);
    uint32_t* s32 = reinterpret_cast<uint32_t*>(s_row);
    for (int row = blockIdx.x; row < n_edge + n_var; row += gridDim.x) {
        uint8_t* base = row < n_edge ? msg + (size_t)row * pitch : ch + (size_t)(row - n_edge) * pitch;
        __syncthreads();   // the previous row has left the staging buffer
        for (int i = threadIdx.x; i < act16; i += kGatherThreads) reinterpret_cast<uint4*>(s_row)[i] = reinterpret_cast<const uint4*>(base)[i];
        __syncthreads();
        for (int i = threadIdx.x; i < n_holes; i += kGatherThreads) {
            const int d = hole_dst[i], c = hole_src[i];
            const uint32_t nib = (s32[c >> 3] >> (4 * (c & 7))) & 15u;            // tail columns are never patched
            const uint32_t old = (s32[d >> 3] >> (4 * (d & 7))) & 15u;            // only this thread changes nibble d
            atomicXor(&s32[d >> 3], (old ^ nib) << (4 * (d & 7)));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < new16; i += kGatherThreads) reinterpret_cast<uint4*>(base)[i] = reinterpret_cast<const uint4*>(s_row)[i];
    }
}

// New column layout: all alive; the original frame index of every moved column follows it.  The last block flips the state.
__global__ void pf_commit_kernel(PfState* st, uint32_t* alive, uint32_t* fsyn, int* idx, const int* hole_dst, const int* hole_src)
{
    if (st->done || !st->do_compact) return;
    const int new_n = st->new_n, n_holes = st->n_holes;
    const int old_words = (st->n_act + 7) >> 3;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < old_words; w += gridDim.x * blockDim.x) {
        const int nv = new_n - 8 * w;
        alive[w] = nv >= 8 ? 0xffffffffu : nv <= 0 ? 0u : ((1u << (4 * nv)) - 1u);
        fsyn[w] = 0u;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_holes; i += gridDim.x * blockDim.x) idx[hole_dst[i]] = idx[hole_src[i]];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&st->blocks_done2, 1u);
        if (done == gridDim.x - 1) {          // every block has read the old n_act
            st->n_act = new_n;
            st->act_pitch = (((new_n + 1) / 2 + 15) / 16) * 16;
            st->do_compact = 0;
            st->waste = 0.f;
            st->blocks_done2 = 0u;
        }
    }
}

// result nibbles in slot order -> the caller's uint8 output in frame order: out[row][f] = nibble slot_of[f] of res[row].
// `rows_per_pass` result rows are staged in shared memory at a time (0: read them from global memory).
__global__ void __launch_bounds__(kExpandThreads) pf_expand_kernel(const uint8_t* __restrict__ res, uint32_t res_pitch,
                                                                  const int* __restrict__ slot_of, uint8_t* __restrict__ out, int n_var,
                                                                  int B, uint32_t out_pitch, int rows_per_pass)
{
    extern __shared__ __align__(16) uint8_t s_res[];
    const int words = (B + 7) >> 3;
    const int rp = rows_per_pass > 0 ? rows_per_pass : 1;
    for (int r0 = blockIdx.x * rp; r0 < n_var; r0 += gridDim.x * rp) {
        const int nr = min(rp, n_var - r0);
        if (rows_per_pass > 0) {
            __syncthreads();
            const int n16 = (int)(res_pitch >> 4) * nr;
            const uint4* src = reinterpret_cast<const uint4*>(res + (size_t)r0 * res_pitch);
            for (int i = threadIdx.x; i < n16; i += kExpandThreads) reinterpret_cast<uint4*>(s_res)[i] = src[i];
            __syncthreads();
        }
        for (int w = threadIdx.x; w < words; w += kExpandThreads) {
            int sl[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) sl[q] = 8 * w + q < B ? slot_of[8 * w + q] : 0;
            for (int r = 0; r < nr; ++r) {
                const uint8_t* row = rows_per_pass > 0 ? s_res + (size_t)r * res_pitch : res + (size_t)(r0 + r) * res_pitch;
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    lo |= ((uint32_t)(row[sl[q] >> 1] >> (4 * (sl[q] & 1))) & 15u) << (8 * q);
                    hi |= ((uint32_t)(row[sl[q + 4] >> 1] >> (4 * (sl[q + 4] & 1))) & 15u) << (8 * q);
                }
                uint8_t* dst = out + (size_t)(r0 + r) * out_pitch + 8 * (size_t)w;
                if (8 * (uint32_t)w + 8 <= out_pitch && (out_pitch & 7u) == 0u) {
                    *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
                } else {
                    for (int f = 0; f < 8; ++f)
                        if (8 * (uint32_t)w + f < out_pitch) dst[f] = (uint8_t)(((f < 4 ? lo : hi) >> (8 * (f & 3))) & 0xffu);
                }
            }
        }
    }
}

int ensure_buf(void** p, size_t* have, size_t need)
{
    if (*have >= need && *p) return IBLDPC_OK;
    if (*p) IBLDPC_CK(cudaFree(*p));
    *p = nullptr;
    *have = 0;
    if (cudaMalloc(p, need) != cudaSuccess) return fail_msg(IBLDPC_E_NOMEM, "cudaMalloc of " + std::to_string(need) + " bytes failed");
    *have = need;
    return IBLDPC_OK;
}

}  // namespace

// `a`: graph pointers, a.ch = packed channel values (w.ch4), a.msg = w.msg, a.out = caller's uint8 output, pitches, T.
// i_num_frames_dev: int32 [B] on the device or nullptr.
int decode_ib_perframe(ibldpc_decoder* h, Workspace& w, const IbArgs& a, long long B, int imax, int32_t* i_num_frames_dev,
                       cudaStream_t st)
{
    const PhaseSetOps* ops = phase_ops_of(h);
    if (!ops) return fail_msg(IBLDPC_E_STATE, "per-frame early termination runs on the fused per-phase kernels: this code's degree sets are not instantiated (or IBLDPC_NO_PHASE is set)");
    const uint32_t pitch4 = a.pitch;
    const int words = (int)(pitch4 / 4);
    int rc;
    if ((rc = phase_set_attributes(h))) return rc;
    // result slots: one per frame + at most 7 padding slots per group (imax groups)
    const size_t slots = (size_t)words * 8 + 8 * (size_t)(imax + 1);
    const uint32_t res_pitch = (uint32_t)(((slots / 2 + 15) / 16) * 16);
    {
        // ints: [idx][hole_dst][hole_src][slot_of] of words*8, [fin] of slots, [gstart] of imax+2 (rounded to 4);
        // words: [alive][fsyn]; state (256 B); [result nibbles n_var x res_pitch]
        const size_t n_int = (size_t)words * 8 * 4 + slots + (size_t)((imax + 2 + 3) / 4 * 4);
        const size_t need = sizeof(int) * n_int + sizeof(uint32_t) * (size_t)words * 2 + 256 + 16 + (size_t)h->N * res_pitch;
        void* p = w.pf_idx;
        if ((rc = ensure_buf(&p, &w.pf_idx_bytes, need))) return rc;
        w.pf_idx = (int*)p;
    }
    int* idx0 = w.pf_idx;
    int* hole_dst = idx0 + (size_t)words * 8;
    int* hole_src = hole_dst + (size_t)words * 8;
    int* slot_of = hole_src + (size_t)words * 8;
    int* fin = slot_of + (size_t)words * 8;
    int* gstart = fin + slots;
    uint32_t* alive = reinterpret_cast<uint32_t*>(gstart + (size_t)((imax + 2 + 3) / 4 * 4));
    uint32_t* fsyn = alive + words;
    PfState* state = reinterpret_cast<PfState*>(fsyn + words);
    uint8_t* res = reinterpret_cast<uint8_t*>(fsyn + words) + 256;
    res += (16 - (reinterpret_cast<uintptr_t>(res) & 15)) & 15;
    static_assert(sizeof(PfState) <= 256, "PfState outgrew its slot");
    uint8_t* msg = w.msg;
    uint8_t* ch4 = w.ch4;
    if (a.msg != msg || a.ch != ch4) return fail_msg(IBLDPC_E_STATE, "per-frame early termination: unexpected workspace buffers");

    const int small_grid = std::max(1, std::min(h->sm_count * 4, (words + 255) / 256));
    pf_init_kernel<<<small_grid, 256, 0, st>>>(state, (int)B, (int)pitch4, alive, fsyn, idx0, gstart, words);
    h->last_launches++;

    PhaseArgs base{};
    base.a = a;
    base.a.early = 1;
    base.a.imax = imax;
    base.pf = state;
    base.pf_alive = alive;
    base.pf_fsyn = fsyn;
    base.pf_fin = fin;
    base.pf_gstart = gstart;
    base.pf_res = res;
    base.pf_res_pitch = res_pitch;
    auto launch = [&](int mode, int index, int it) -> int {
        PhaseArgs q = base;
        q.a.it = it;
        q.a.iter0 = (mode == kPhaseCn && index == 0 && it < 0);
        size_t smem = 0;
        phase_fill_args(h, mode, index, q, &smem);
        PhaseKernel k = ops->vn_pf_kernel;
        if (mode == kPhaseCn) {
            // syndrome accumulator behind the image when the whole batch fits (227 KB per CTA minus image and statics)
            const size_t room = (size_t)227 * 1024 - 1024 - smem;
            const bool fits = (size_t)words * 4 <= room;
            if (fits) smem += (size_t)words * 4;
            k = ops->cn_pf_kernel[fits ? 0 : 1];
        }
        k<<<h->sm_count, kPhaseThreads, smem, st>>>(q);
        h->last_launches++;
        return IBLDPC_OK;
    };
    auto retire = [&](int it, int last) {
        pf_update_kernel<<<small_grid, 256, 0, st>>>(state, alive, fsyn, idx0, fin, slot_of, gstart, i_num_frames_dev, a.inum, it, last, imax);
        h->last_launches++;
    };
    // decision of the groups [g_lo, g_hi] (group g = pass g - 1, tables of iteration g)
    int g_next = 0;
    auto decide_pending = [&](int g_hi) {
        if (g_hi < g_next) return;
        PhaseArgs q = base;
        size_t smem = 0;
        phase_fill_args(h, kPhaseOut, 0, q, &smem);
        ops->pf_decide_kernel<<<h->sm_count, kPhaseThreads, smem, st>>>(q, g_next, g_hi);
        h->last_launches++;
        g_next = g_hi + 1;
    };
    // cost of one compaction in passes of the same width (measured on B200: gather + scan + commit against a
    // variable-node + check-node pass of the (3,6) and 802.11n codes); IBLDPC_PF_COST overrides
    static const float cost = getenv("IBLDPC_PF_COST") ? (float)atof(getenv("IBLDPC_PF_COST")) : 0.3f;
    const size_t row_smem = std::min<size_t>(pitch4, kRowSmemMax);
    if (!w.pf_attr_set) {
        IBLDPC_CK(cudaFuncSetAttribute((const void*)pf_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowSmemMax));
        IBLDPC_CK(cudaFuncSetAttribute((const void*)pf_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRowSmemMax));
        w.pf_attr_set = true;
    }
    const int gather_grid = h->sm_count * (int)std::max<size_t>(1, std::min<size_t>(4, kRowSmemMax / std::max<size_t>(row_smem, 1)));
    auto compact = [&](int it) {
        pf_scan_kernel<<<1, 1024, 0, st>>>(state, alive, hole_dst, hole_src, imax - 2 - it, cost, (int)kRowSmemMax);
        decide_pending(it + 1);          // the frozen columns are read before the gather may overwrite them
        pf_gather_kernel<<<gather_grid, kGatherThreads, row_smem, st>>>(state, msg, ch4, h->E, h->N, pitch4, hole_dst, hole_src);
        pf_commit_kernel<<<small_grid, 256, 0, st>>>(state, alive, fsyn, idx0, hole_dst, hole_src);
        h->last_launches += 3;
    };
    // check-node phase of iteration 0 (table block 0, channel values through vidx)
    if ((rc = launch(kPhaseCn, 0, -1))) return rc;
    if (imax <= 1) retire(-1, 1);          // no pass at all: decide everything with table 0, i_num = 1
    for (int it = 0; it < imax - 1; ++it) {
        if ((rc = launch(kPhaseVn, it, it))) return rc;
        if ((rc = launch(kPhaseCn, it + 1, it))) return rc;
        retire(it, it == imax - 2);
        // compaction attempts: every pass in the first half of the schedule (the waterfall region retires most frames
        // within a few passes), then every second / fourth pass; whether one happens is decided on the device
        const bool attempt = it < imax - 2 && (it < 24 || (it < 36 && it % 2 == 1) || it % 4 == 3);
        if (attempt) compact(it);
    }
    decide_pending(std::max(imax - 1, 0));
    {
        const int rows_per_pass = res_pitch <= kRowSmemMax ? (int)std::min<size_t>(4, kRowSmemMax / res_pitch) : 0;
        const size_t smem = (size_t)rows_per_pass * res_pitch;
        const int g = std::max(1, std::min(h->sm_count * (smem <= 100 * 1024 ? 2 : 1), (h->N + std::max(rows_per_pass, 1) - 1) / std::max(rows_per_pass, 1)));
        pf_expand_kernel<<<g, kExpandThreads, smem, st>>>(res, res_pitch, slot_of, a.out, h->N, (int)B, a.out_pitch, rows_per_pass);
        h->last_launches++;
    }
    IBLDPC_CK(cudaGetLastError());
    return IBLDPC_OK;
}

}  // namespace ibldpc
