// ibldpc_internal.h -- state shared by the translation units of libibldpc.so (not part of the C ABI).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ibldpc.h"

namespace ibldpc {

// sets the calling thread's last-error string (ibldpc_last_error) and returns `code`
int fail_msg(int code, const std::string& msg);

// Makes `device` current for the duration of a C-ABI call and restores the caller's device afterwards
// (one handle per GPU, but a host process may drive several GPUs from one thread).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = err == cudaSuccess;
        }
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct NodeClass {
    int degree = 0;
    int count = 0;
    int* d_nodes = nullptr;
};

struct PhaseEvent { cudaEvent_t a, b; int phase; };

constexpr int kMaxIter = 4096;

struct Workspace {
    uint8_t* msg = nullptr;     // IB in-place message array
    size_t msg_bytes = 0;
    uint8_t* ch4 = nullptr;     // packed-nibble copy of the channel values (n4 path)
    size_t ch4_bytes = 0;
    void* cin = nullptr;        // LLR inboxes
    void* vin = nullptr;
    size_t llr_bytes = 0;
    uint8_t* padbuf_in = nullptr;   // padded copies of caller buffers when B is not vector-aligned
    uint8_t* padbuf_out = nullptr;
    size_t pad_bytes = 0;
    uint8_t* stage_in = nullptr;    // device staging of the host-buffer path
    uint8_t* stage_out = nullptr;
    size_t stage_bytes = 0;
    uint8_t* stage_bits = nullptr;  // bit-packed hard decisions of the packed host path
    size_t stage_bits_bytes = 0;
    int* flags = nullptr;           // [kMaxIter] batch syndrome flags + [kMaxIter] = input range flag
    int* inum = nullptr;
    cudaStream_t stream = nullptr;  // host-path stream
    // degree classes of one phase run concurrently on these (fork/join around every phase)
    cudaStream_t aux[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr, join_ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // per-frame early termination (in-place frame compaction, deferred decisions): see ib_perframe.cu
    int* pf_idx = nullptr;          // index lists, masks, device-side state and the dense result nibbles
    size_t pf_idx_bytes = 0;
    bool pf_attr_set = false;
    // pinned host staging of the int32 host contract (two slots)
    uint8_t* pin_in = nullptr;
    uint8_t* pin_out = nullptr;
    size_t pin_bytes = 0;
    cudaEvent_t done_ev = nullptr;
};

struct PhaseImages;   // ib_phase.cu: pre-expanded shared-memory table images of the fused per-phase kernels
struct T32Images;     // ib_t32.cu: table images of the |T| <= 32 family
struct LayerPlan;     // llr_layered.cu: check-node layers of the layered LLR schedule
struct IbArgs;

}  // namespace ibldpc

struct ibldpc_decoder {
    int device = 0;
    int sm_count = 148;
    int N = 0, M = 0, E = 0;
    int dc_max = 0, dv_max = 0, dc_min = 0, dv_min = 0;
    int *d_sc = nullptr, *d_dc = nullptr, *d_tc = nullptr, *d_sv = nullptr, *d_dv = nullptr, *d_tv = nullptr,
        *d_vidx = nullptr;
    std::vector<int> h_sc, h_dc, h_sv, h_dv, h_tv, h_vidx;   // host copies (work lists of the fused per-phase kernels, layer colouring)
    std::vector<ibldpc::NodeClass> cn_classes, vn_classes;
    // LUTs
    bool have_luts = false;
    int T = 0, Tc = 0, lut_imax = 0, DC = 0, DV = 0;
    bool match = false;
    uint8_t *d_cn8 = nullptr, *d_vn8 = nullptr, *d_mc8 = nullptr, *d_mv8 = nullptr;
    std::vector<uint8_t> h_cn8, h_vn8, h_mc8, h_mv8;   // host copies of the uint8 tables (image builders)
    uint8_t* d_cn_pair = nullptr;   // [imax blocks][cn classes][T*T rows][8 bytes] composed tail-pair tables
    uint8_t* d_vn_pair = nullptr;   // [imax][vn classes][T*T rows][8 bytes] composed tail-pair tables of the VN update
    std::vector<uint8_t> h_cn3, h_vn3;   // host copies of d_cn3 / d_vn3 (phase images of the three-input-table set)
    uint8_t* d_vn3 = nullptr;       // [imax][16*16*16] three-input table of the degree-3 variable-node update (ib_triple_n4.cuh)
    uint8_t* d_cn3 = nullptr;       // [imax blocks][16*16*16] three-input table of the first two check-node stages, values * 4
    int cn_tri_max_degree = 8;      // check-node classes of degree 6..this run the three-input-table kernel (IBLDPC_CN_TRI_MAX_DEGREE)
    bool use_triple = true;         // IBLDPC_NO_TRIPLE=1: degree-3 variable nodes through the two-input stage tables
    std::vector<uint8_t> h_cn_pair, h_vn_pair;
    int vn_pair_min_degree = 5;     // packed-nibble family (IBLDPC_VN_PAIR_MIN_DEGREE)
    int vn_pair_threads = 0;        // 0 = per-degree default, 256 / 512 forced (IBLDPC_VN_PAIR_THREADS)
    int cn_threads = 0, vn_threads = 0;   // 0 = default CTA sizes (1024 where instantiated); IBLDPC_CN_THREADS=512 /
                                          // IBLDPC_VN_THREADS=256 select the smaller CTAs (parity variants, A/B)
    long long coop_max_frames = 4096;   // whole-decode cooperative kernel up to this batch size (policy: end of ibldpc_set_luts)
    long long phase_mid_max_frames = 4096;   // instantiated sets without phase_default: fused per-phase kernels up to this batch size
    bool phase_off_midrange = false;    // IBLDPC_NO_PHASE
    bool coop_max_from_env = false;     // IBLDPC_COOP_MAX_B given: no message-size limit on the cooperative phase-image kernel
    bool no_coop_phase = false;         // IBLDPC_NO_COOP_PHASE=1: small batches through the table-restaging cooperative kernels (ib_coop_n4.cuh)
                                        // (IBLDPC_COOP_MAX_B, 0 disables)
    int coop_supported = -1;        // device attribute cudaDevAttrCooperativeLaunch, queried once
    bool use_pair = true;
    int pair_min_degree = 7;      // uint8 family
    int n4_pair_min_degree = 6;   // packed-nibble family
    bool fast = false;
    bool nib = false;     // packed-nibble fast path (ib_kernels_n4.cuh)
    bool t32 = false;     // |T| in (16, 32]: shared-memory byte family (ib_kernels_t32.cuh)
    int vn_vec = 0;       // words per lane of the variable-node kernels: 0 = per-degree default, 2 / 4 forced (IBLDPC_VN_VEC)
    int Wc = 1, Wv = 1, Wo = 1, nrows_c = 0, nrows_v = 0, nrows_o = 0, tshift = -1;
    ibldpc::Workspace ws[2];
    int host_chunk = 0;   // 0 = auto: about 256 MiB of channel values per chunk
    // fused per-phase kernels (ib_phase_n4.cuh): one launch per phase over all degree classes
    ibldpc::PhaseImages* phase = nullptr;
    ibldpc::PhaseImages* phase_tri = nullptr;   // (3,6) set through the three-input tables: per-frame early termination runs on it
    ibldpc::T32Images* t32_images = nullptr;
    ibldpc::LayerPlan* layers = nullptr;   // layered LLR schedule (llr_layered.cu), built at the first layered decode
    int use_phase = 1;    // IBLDPC_NO_PHASE=1 keeps one launch per degree class
    bool phase_default = false;   // plain decodes use the fused kernels (802.11n sets; IBLDPC_PHASE=1: every instantiated set)
    // request of ibldpc_decode_ib_perframe for the decode being issued
    bool pf_request = false;
    int32_t* pf_inum = nullptr;
    // stream / workspace of the last decode (lazy i_num read-back)
    cudaStream_t last_stream = nullptr;
    int last_ws = 0;
    // introspection
    int last_launches = 0, last_grid = 0, last_smem = 0;
    bool profiling = false;
    std::vector<ibldpc::PhaseEvent> events;
    std::map<std::pair<const void*, int>, int> occ_cache;
    // NCCL communicator of the counter all-reduce (nccl_abi.cu), opaque here
    void* nccl_comm = nullptr;
};

namespace ibldpc {
// fused per-phase kernels (ib_phase.cu)
int phase_prepare(ibldpc_decoder* h);   // end of ibldpc_set_luts: build + upload the phase images (or leave h->phase null)
void phase_free(ibldpc_decoder* h);
int decode_ib_phase(ibldpc_decoder* h, const IbArgs& a, int imax, int early, cudaStream_t st);
// B <= kLaneModeMaxFrames: the whole decode in one cooperative launch over the same phase images (ib_coop_phase_kernel)
int decode_ib_coop_phase(ibldpc_decoder* h, const IbArgs& a, long long B, int imax, int early, cudaStream_t st);
// per-frame early termination with frame compaction (ib_perframe.cu)
int decode_ib_perframe(ibldpc_decoder* h, Workspace& w, const IbArgs& a, long long B, int imax, int32_t* i_num_frames_dev,
                       cudaStream_t st);
// |T| <= 32 shared-memory family (ib_t32.cu)
int t32_prepare(ibldpc_decoder* h);
void t32_free(ibldpc_decoder* h);
int decode_ib_t32(ibldpc_decoder* h, const IbArgs& a, int imax, int early, cudaStream_t st);
// layered LLR schedule (llr_layered.cu)
void layered_free(ibldpc_decoder* h);
int layered_prepare(ibldpc_decoder* h);          // colour the checks (once per handle)
int layered_count(const ibldpc_decoder* h);
int decode_llr_layered_f32(ibldpc_decoder* h, Workspace& w, int algo, const float* ch, long long pitch, long long B, int imax,
                           int early, float* out, cudaStream_t st);
int decode_llr_layered_f64(ibldpc_decoder* h, Workspace& w, int algo, const double* ch, long long pitch, long long B, int imax,
                           int early, double* out, cudaStream_t st);
}  // namespace ibldpc

#define IBLDPC_CK(call)                                                                                       \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return ::ibldpc::fail_msg(IBLDPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)
