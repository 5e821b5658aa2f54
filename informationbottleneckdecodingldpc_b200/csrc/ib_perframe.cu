// ib_perframe.cu -- opt-in PER-FRAME early termination with frame compaction.
//
// The reference stops a whole batch at once (discrete_LDPC_decoder.py:233,273: `while i_num < imax and not
// syndrome_zero` on the SUM of the syndrome over all msg_at_time frames).  Decoding every frame on its own
// (msg_at_time = 1) stops each frame as soon as ITS syndrome is zero; that is the result this mode reproduces for a
// whole batch in one call -- outputs and per-frame i_num equal to the reference run with one frame per call -- while
// only the frames that still iterate cost anything:
//   * every check-node phase records, per frame, whether a check failed (bit 4f of pf_fsyn[word], atomicOr);
//   * pf_update_kernel turns "alive and no failed check in this pass" into the nibble mask pf_conv, stores the frame's
//     i_num, and counts what is left;
//   * the decision kernel of the SAME pass (table of iteration i_num - 1, like calc_varnode_output at
//     discrete_LDPC_decoder.py:280-287) writes exactly those frames, to their original columns;
//   * at scheduled passes the surviving columns are gathered to the front of the other ping-pong buffer (packed
//     nibbles: one word-gather kernel) when at least a quarter of the active columns has finished, and every later
//     kernel works on the shorter prefix.
// Nothing synchronises with the host: all launches of the schedule are issued up front and read what is left to do
// from a device-side PfState.  Runs on the fused per-phase kernels (ib_phase_n4.cuh), i.e. for the degree sets those
// are instantiated for.
#include <algorithm>
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

__global__ void pf_init_kernel(PfState* st, int B, int pitch4, uint32_t* alive, uint32_t* fsyn, uint32_t* conv, int* idx0,
                               int* dstw, int words)
{
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        dstw[w] = w;
        const int nv = B - 8 * w;
        alive[w] = nv >= 8 ? 0xffffffffu : nv <= 0 ? 0u : ((1u << (4 * nv)) - 1u);
        fsyn[w] = 0u;
        conv[w] = 0u;
    }
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < words * 8; f += gridDim.x * blockDim.x) idx0[f] = f < B ? f : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->n_act = B; st->act_pitch = pitch4; st->cur = 0; st->n_alive = B; st->done = 0; st->do_compact = 0;
        st->new_n = B; st->alive_acc = 0; st->blocks_done = 0u;
    }
}

// After the check-node phase of pass `it`: conv = alive frames without a failed check (all alive frames in the last
// pass), i_num of those frames = it + 2 (passes executed + 1), alive -= conv, fsyn = 0, count the rest.
__global__ void pf_update_kernel(PfState* st, uint32_t* alive, uint32_t* fsyn, uint32_t* conv, const int* idx0, const int* idx1,
                                 int32_t* inum_frames, int* inum_batch, int it, int last)
{
    if (st->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) *inum_batch = it + 2;   // i_num of the last frame to finish = the batch's i_num
    const int words = (st->n_act + 7) >> 3;
    const int* idx = st->cur ? idx1 : idx0;
    int cnt = 0;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
        const uint32_t al = alive[w];
        const uint32_t failed = (fsyn[w] & 0x11111111u) * 15u;
        const uint32_t cv = last ? al : (al & ~failed);
        conv[w] = cv;
        alive[w] = al & ~cv;
        fsyn[w] = 0u;
        cnt += __popc(al & ~cv) >> 2;
        if (cv != 0u && inum_frames != nullptr) {
#pragma unroll
            for (int f = 0; f < 8; ++f)
                if ((cv >> (4 * f)) & 1u) inum_frames[idx[8 * w + f]] = it + 2;
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
            // st->done is raised by pf_finish_kernel AFTER the decision kernel of this pass has written its frames
        }
    }
}

__global__ void pf_finish_kernel(PfState* st)
{
    if (st->n_alive == 0) st->done = 1;
}

// One CTA: if at least a quarter of the active columns has finished, list the alive columns (stable order) and the
// original frame index of each; otherwise leave do_compact = 0 and the gather is skipped.
__global__ void __launch_bounds__(1024) pf_scan_kernel(PfState* st, const uint32_t* alive, const int* idx0, const int* idx1,
                                                       int* order, int* idx0w, int* idx1w)
{
    __shared__ int s_sum[1024];
    if (st->done) return;
    const int n_act = st->n_act, n_alive = st->n_alive;
    if (n_alive == 0 || (long long)n_alive * 4 > (long long)n_act * 3) {
        if (threadIdx.x == 0) st->do_compact = 0;
        return;
    }
    const int words = (n_act + 7) >> 3;
    const int per = (words + 1023) / 1024;
    const int w0 = threadIdx.x * per, w1 = min(words, w0 + per);
    int cnt = 0;
    for (int w = w0; w < w1; ++w) cnt += __popc(alive[w]) >> 2;
    s_sum[threadIdx.x] = cnt;
    __syncthreads();
    // inclusive scan (Hillis-Steele, 1024 entries)
    for (int off = 1; off < 1024; off <<= 1) {
        const int v = threadIdx.x >= off ? s_sum[threadIdx.x - off] : 0;
        __syncthreads();
        s_sum[threadIdx.x] += v;
        __syncthreads();
    }
    int pos = s_sum[threadIdx.x] - cnt;
    const int* idx = st->cur ? idx1 : idx0;
    int* idxw = st->cur ? idx0w : idx1w;          // the OTHER buffer receives the compacted index list
    for (int w = w0; w < w1; ++w) {
        const uint32_t al = alive[w];
        for (int f = 0; f < 8; ++f)
            if ((al >> (4 * f)) & 1u) {
                order[pos] = 8 * w + f;
                idxw[pos] = idx[8 * w + f];
                ++pos;
            }
    }
    if (threadIdx.x == 1023) {
        st->new_n = s_sum[1023];
        st->do_compact = 1;
    }
}

// dst[row][word j] = nibbles of the alive columns order[8j .. 8j+7] of src[row]; rows = message rows, then channel rows
__global__ void pf_gather_kernel(const PfState* st, uint8_t* msg0, uint8_t* msg1, uint8_t* ch0, uint8_t* ch1, int n_edge,
                                 int n_var, uint32_t pitch, const int* __restrict__ order)
{
    if (st->done || !st->do_compact) return;
    const int new_n = st->new_n;
    const int nw = (new_n + 7) >> 3;
    const uint8_t* smsg = st->cur ? msg1 : msg0;
    uint8_t* dmsg = st->cur ? msg0 : msg1;
    const uint8_t* sch = st->cur ? ch1 : ch0;
    uint8_t* dch = st->cur ? ch0 : ch1;
    const long long total = (long long)(n_edge + n_var) * nw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / nw), j = (int)(i - (long long)row * nw);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(row < n_edge ? smsg + (size_t)row * pitch : sch + (size_t)(row - n_edge) * pitch);
        uint32_t* dst = reinterpret_cast<uint32_t*>(row < n_edge ? dmsg + (size_t)row * pitch : dch + (size_t)(row - n_edge) * pitch);
        uint32_t v = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int c = 8 * j + q < new_n ? order[8 * j + q] : -1;
            if (c >= 0) v |= ((src[c >> 3] >> (4 * (c & 7))) & 15u) << (4 * q);
        }
        dst[j] = v;
    }
}

__global__ void pf_commit_kernel(PfState* st, uint32_t* alive, uint32_t* fsyn, uint32_t* conv, int old_words_max)
{
    if (st->done || !st->do_compact) return;
    const int new_n = st->new_n;
    const int old_words = (st->n_act + 7) >> 3;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < old_words && w < old_words_max; w += gridDim.x * blockDim.x) {
        const int nv = new_n - 8 * w;
        alive[w] = nv >= 8 ? 0xffffffffu : nv <= 0 ? 0u : ((1u << (4 * nv)) - 1u);
        fsyn[w] = 0u;
        conv[w] = 0u;
    }
    // grid-wide: the state flips after every block has read the old n_act -- done by a second, one-thread launch
}

// destination-word shortcut of the decision kernel for the compacted order (reads the NEW index list)
__global__ void pf_dstw_kernel(const PfState* st, const int* idx0, const int* idx1, int* dstw)
{
    if (st->done || !st->do_compact) return;
    const int* idx = st->cur ? idx0 : idx1;          // the buffer the scan wrote = the one that becomes current
    const int nw = (st->new_n + 7) >> 3;
    for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += gridDim.x * blockDim.x) {
        const int i0 = idx[8 * w], i7 = idx[8 * w + 7];
        dstw[w] = (8 * w + 8 <= st->new_n && (i0 & 7) == 0 && i7 - i0 == 7) ? (i0 >> 3) : -1;
    }
}

__global__ void pf_flip_kernel(PfState* st)
{
    if (st->done || !st->do_compact) return;
    st->n_act = st->new_n;
    st->act_pitch = (((st->new_n + 1) / 2 + 15) / 16) * 16;
    st->cur ^= 1;
    st->do_compact = 0;
}

// result nibbles (original frame order) -> the caller's uint8 output
__global__ void pf_expand_kernel(const uint8_t* __restrict__ res, uint8_t* __restrict__ out, int n_var, int B, uint32_t pitch4,
                                 uint32_t out_pitch)
{
    const uint32_t wpr = pitch4 >> 2;
    const long long total = (long long)n_var * wpr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / wpr), w = (int)(i - (long long)row * wpr);
        const uint32_t v = reinterpret_cast<const uint32_t*>(res + (size_t)row * pitch4)[w];
        uint8_t* dst = out + (size_t)row * out_pitch + 8 * (size_t)w;
        if (8 * w + 8 <= (int)out_pitch && (out_pitch & 7u) == 0u) {
            const uint32_t lo = (v & 0xfu) | ((v & 0xf0u) << 4) | ((v & 0xf00u) << 8) | ((v & 0xf000u) << 12);
            const uint32_t hv = v >> 16;
            const uint32_t hi = (hv & 0xfu) | ((hv & 0xf0u) << 4) | ((hv & 0xf00u) << 8) | ((hv & 0xf000u) << 12);
            *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
        } else {
            for (int f = 0; f < 8; ++f)
                if (8 * w + f < (int)out_pitch) dst[f] = (uint8_t)((v >> (4 * f)) & 15u);
        }
    }
    (void)B;
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
    // second message / channel arrays, index lists, masks, state
    {
        void* p = w.pf_msg2;
        if ((rc = ensure_buf(&p, &w.pf_msg2_bytes, (size_t)h->E * pitch4))) return rc;
        w.pf_msg2 = (uint8_t*)p;
        p = w.pf_ch2;
        if ((rc = ensure_buf(&p, &w.pf_ch2_bytes, (size_t)h->N * pitch4))) return rc;
        w.pf_ch2 = (uint8_t*)p;
        // [idx0][idx1][order] ints of words*8, [alive][fsyn][conv] words, state, [result nibbles n_var x pitch4]
        const size_t need = sizeof(int) * (size_t)words * 8 * 3 + sizeof(uint32_t) * (size_t)words * 4 + 256 + (size_t)h->N * pitch4;
        p = w.pf_idx;
        if ((rc = ensure_buf(&p, &w.pf_idx_bytes, need))) return rc;
        w.pf_idx = (int*)p;
    }
    int* idx0 = w.pf_idx;
    int* idx1 = idx0 + (size_t)words * 8;
    int* order = idx1 + (size_t)words * 8;
    uint32_t* alive = reinterpret_cast<uint32_t*>(order + (size_t)words * 8);
    uint32_t* fsyn = alive + words;
    uint32_t* conv = fsyn + words;
    int* dstw = reinterpret_cast<int*>(conv + words);
    PfState* state = reinterpret_cast<PfState*>(dstw + words);
    uint8_t* res = reinterpret_cast<uint8_t*>(dstw + words) + 256;
    IBLDPC_CK(cudaMemsetAsync(res, 0, (size_t)h->N * pitch4, st));
    const int small_grid = std::max(1, std::min(h->sm_count * 4, (words + 255) / 256));
    pf_init_kernel<<<small_grid, 256, 0, st>>>(state, (int)B, (int)pitch4, alive, fsyn, conv, idx0, dstw, words);
    h->last_launches++;

    PhaseArgs base{};
    base.a = a;
    base.a.early = 1;
    base.a.imax = imax;
    base.pf = state;
    base.pf_msg[0] = a.msg; base.pf_msg[1] = w.pf_msg2;
    base.pf_ch[0] = a.ch; base.pf_ch[1] = w.pf_ch2;
    base.pf_idx[0] = idx0; base.pf_idx[1] = idx1;
    base.pf_fsyn = fsyn;
    base.pf_conv = conv;
    base.pf_res = res;
    base.pf_dstw = dstw;
    auto launch = [&](int mode, int index, int it, PhaseKernel k) -> int {
        PhaseArgs q = base;
        q.a.it = it;
        q.a.iter0 = (mode == kPhaseCn && index == 0 && it < 0);
        size_t smem = 0;
        phase_fill_args(h, mode, index, q, &smem);
        if (mode == kPhaseCn) {
            // syndrome accumulator behind the image when the whole batch fits (227 KB per CTA minus image and statics)
            const size_t room = (size_t)227 * 1024 - 1024 - smem;
            if ((size_t)words * 4 <= room) {
                q.pf_fsyn_smem_words = words;
                smem += (size_t)words * 4;
            }
        }
        k<<<h->sm_count, kPhaseThreads, smem, st>>>(q);
        h->last_launches++;
        return IBLDPC_OK;
    };
    auto retire = [&](int it, int last) -> int {
        pf_update_kernel<<<small_grid, 256, 0, st>>>(state, alive, fsyn, conv, idx0, idx1, i_num_frames_dev, a.inum, it, last);
        // decision with the variable-node tables of iteration it + 1 for the frames named by conv
        if ((rc = launch(kPhaseOut, it + 1, it + 1, ops->out_pf_kernel))) return rc;
        pf_finish_kernel<<<1, 1, 0, st>>>(state);
        h->last_launches += 2;
        return IBLDPC_OK;
    };
    auto compact = [&]() -> int {
        pf_scan_kernel<<<1, 1024, 0, st>>>(state, alive, idx0, idx1, order, idx0, idx1);
        const long long total = (long long)(h->E + h->N) * words;
        const int g = (int)std::max<long long>(1, std::min<long long>((long long)h->sm_count * 16, (total + 255) / 256));
        pf_gather_kernel<<<g, 256, 0, st>>>(state, a.msg, w.pf_msg2, const_cast<uint8_t*>(a.ch), w.pf_ch2, h->E, h->N, pitch4, order);
        pf_commit_kernel<<<small_grid, 256, 0, st>>>(state, alive, fsyn, conv, words);
        pf_dstw_kernel<<<small_grid, 256, 0, st>>>(state, idx0, idx1, dstw);
        pf_flip_kernel<<<1, 1, 0, st>>>(state);
        h->last_launches += 5;
        return IBLDPC_OK;
    };
    // check-node phase of iteration 0 (table block 0, channel values through vidx)
    if ((rc = launch(kPhaseCn, 0, -1, ops->cn_pf_kernel))) return rc;
    if (imax <= 1) {
        if ((rc = retire(-1, 1))) return rc;          // no pass at all: decide everything with table 0, i_num = 1
    }
    for (int it = 0; it < imax - 1; ++it) {
        if ((rc = launch(kPhaseVn, it, it, ops->vn_pf_kernel))) return rc;
        if ((rc = launch(kPhaseCn, it + 1, it, ops->cn_pf_kernel))) return rc;
        if ((rc = retire(it, it == imax - 2))) return rc;
        // compaction attempts: every pass at first (the waterfall region retires most frames within a few passes),
        // then every second / fourth pass
        const bool try_compact = it < imax - 2 && (it < 8 || (it < 24 && it % 2 == 1) || it % 4 == 3);
        if (try_compact && (rc = compact())) return rc;
    }
    {
        const long long total = (long long)h->N * words;
        const int g = (int)std::max<long long>(1, std::min<long long>((long long)h->sm_count * 16, (total + 255) / 256));
        pf_expand_kernel<<<g, 256, 0, st>>>(res, a.out, h->N, (int)B, pitch4, a.out_pitch);
        h->last_launches++;
    }
    IBLDPC_CK(cudaGetLastError());
    return IBLDPC_OK;
}

}  // namespace ibldpc
