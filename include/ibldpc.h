/*
 * ibldpc.h -- C ABI of the B200-native LDPC decoding engine (libibldpc.so).
 *
 * This is the drop-in boundary for the reference's data-parallel hot path.  Each entry
 * point names the reference interface it replaces (paths relative to the reference
 * repository root).  Plain pointers and sizes only; no torch / CUDA runtime types.
 *
 * Conventions
 *   - every function returns 0 on success, a negative IBLDPC_E_* code otherwise;
 *     ibldpc_last_error() gives the message of the calling thread's last failure;
 *   - buffers are row-major (rows, B) with the frame index fastest, exactly the
 *     (N_var, msg_at_time) C-order arrays of the reference
 *     (Discrete_LDPC_decoding/kernels_template.cl:27 indexes row*msg_at_time+gid2);
 *   - "_dev" pointers are device pointers on the handle's GPU, "_host" pointers host memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); device-buffer calls
 *     are asynchronous on it unless they return a scalar to the host;
 *   - cluster indices travel as uint8 (the reference uses int32; values are < |T| <= 256).
 */
#ifndef IBLDPC_H
#define IBLDPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IBLDPC_OK 0
#define IBLDPC_E_INVALID (-1)  /* bad argument / inconsistent tables */
#define IBLDPC_E_CUDA (-2)     /* CUDA runtime error                 */
#define IBLDPC_E_STATE (-3)    /* call order (e.g. decode before set_luts) */
#define IBLDPC_E_NOMEM (-4)

#define IBLDPC_ALGO_MINSUM 0
#define IBLDPC_ALGO_BP 1
#define IBLDPC_F32 32
#define IBLDPC_F64 64

typedef struct ibldpc_decoder *ibldpc_handle;

/* The six int32 tables the reference uploads in init_OpenCL_decoding
 * (Discrete_LDPC_decoding/discrete_LDPC_decoder_irreg.py:191-205), host pointers. */
typedef struct ibldpc_code_desc {
    int32_t n_var, n_chk, n_edge;
    const int32_t *inbox_start_chk;  /* [n_chk] inbox_memory_start_checknodes  */
    const int32_t *degree_chk;       /* [n_chk] degree_checknode_nr            */
    const int32_t *target_cells_chk; /* [n_edge] target_memory_cells_checknodes */
    const int32_t *inbox_start_var;  /* [n_var] inbox_memory_start_varnodes    */
    const int32_t *degree_var;       /* [n_var] degree_varnode_nr              */
    const int32_t *target_cells_var; /* [n_edge] target_memory_cells_varnodes  */
} ibldpc_code_desc;

/* Look-up tables of the IB decoder in the reference layout:
 * Trellis_checknode_vector_a / Trellis_varnode_vector_a / matching_vector_* as int32,
 * exactly the arrays init_OpenCL_decoding uploads (discrete_LDPC_decoder_irreg.py:207-213).
 * cn_degree / vn_degree are the CN_DEGREE / VN_DEGREE macros (d_c_max / d_v_max,
 * discrete_LDPC_decoder_irreg.py:184).  cn_match == NULL means MATCH false. */
typedef struct ibldpc_lut_desc {
    int32_t card_channel;  /* cardinality_T_channel      */
    int32_t card_decoder;  /* cardinality_T_decoder_ops  */
    int32_t imax;          /* iterations the tables were designed for */
    int32_t cn_degree, vn_degree;
    const int32_t *cn_lut;
    int64_t cn_lut_len;
    const int32_t *vn_lut;
    int64_t vn_lut_len;
    const int32_t *cn_match;
    int64_t cn_match_len;
    const int32_t *vn_match;
    int64_t vn_match_len;
} ibldpc_lut_desc;

/* Replaces the table upload half of init_OpenCL_decoding (discrete_LDPC_decoder.py:132-200,
 * discrete_LDPC_decoder_irreg.py:172-243, min_sum_decoder_irreg.py:167-218,
 * bp_decoder_irreg.py:167-219).  Validates and copies the tables to `device`. */
int ibldpc_create(const ibldpc_code_desc *code, int device, ibldpc_handle *out);

/* Replaces the LUT upload of init_OpenCL_decoding and update_trellis_vectors
 * (discrete_LDPC_decoder.py:53-55).  May be called again to swap tables. */
int ibldpc_set_luts(ibldpc_handle h, const ibldpc_lut_desc *luts);

/* Replaces decode_OpenCL with buffer_in=True, return_buffer=True
 * (discrete_LDPC_decoder.py:202-295, discrete_LDPC_decoder_irreg.py:245-341):
 * send + checknode_update_iter0 + (imax-1) x {varnode_update, checknode_update, calc_syndrome}
 * + calc_varnode_output.  ch_dev / out_dev: uint8 (n_var, B).
 * early_term = 1 reproduces the reference's batch-granular stop (all B frames of the call
 * have zero syndrome); 0 always runs imax-1 passes.  If i_num_host != NULL the call
 * synchronises the stream and stores the reference's i_num (number of passes + 1). */
int ibldpc_decode_ib(ibldpc_handle h, const uint8_t *ch_dev, int64_t B, int imax, int early_term,
                     uint8_t *out_dev, int32_t *i_num_host, void *stream);

/* Replaces decode_OpenCL with buffer_in=False, return_buffer=False (numpy in, numpy out;
 * H2D at discrete_LDPC_decoder_irreg.py:250, D2H at :340) and decode_on_host.
 * Host buffers uint8 (n_var, B); copies are issued inside the call (pinned memory overlaps
 * them with compute when early_term == 0: frames are processed in independent chunks). */
int ibldpc_decode_ib_host(ibldpc_handle h, const uint8_t *ch_host, int64_t B, int imax, int early_term,
                          uint8_t *out_host, int32_t *i_num_host);

/* Opt-in, beyond the reference: PER-FRAME early termination with frame compaction.  The
 * reference stops a batch only when ALL its frames have zero syndrome (discrete_LDPC_decoder.py:233,273); this call
 * returns, for every frame, exactly what the reference returns when that frame is decoded on its own
 * (msg_at_time = 1): decided cluster indices out_dev (n_var, B) and, if i_num_frames_dev != NULL, the frame's own
 * i_num (int32 [B], device).  Converged frames are decided at once and dropped from the message arrays, so the cost
 * follows the average -- not the maximum -- iteration count.  Asynchronous on `stream`; ibldpc_last_i_num gives the
 * largest per-frame i_num.  Needs the packed-nibble family and an instantiated degree set (IBLDPC_E_STATE otherwise). */
int ibldpc_decode_ib_perframe(ibldpc_handle h, const uint8_t *ch_dev, int64_t B, int imax, uint8_t *out_dev,
                              int32_t *i_num_frames_dev, void *stream);

/* i_num of the last decode issued on this handle, read back lazily: synchronises the stream of that decode.  Lets
 * BER loops call ibldpc_decode_ib with i_num_host == NULL (fully asynchronous) and still query the reference's
 * i_num afterwards.  Returns IBLDPC_E_INVALID if the channel buffer of that decode held a value >= card_channel
 * (such values are clamped on the device, never used as table indices). */
int ibldpc_last_i_num(ibldpc_handle h, int32_t *i_num_host);

/* The reference's own host contract (numpy int32 in, numpy int32 out: discrete_LDPC_decoder.py:207-209 casts
 * received_blocks to int32 for the upload, :292-295 returns the int32 output array).  ch_host / out_host: int32
 * (n_var, B).  The narrowing to one byte per cluster index and the widening back run on host threads
 * (IBLDPC_HOST_THREADS, default min(16, cores)) chunk by chunk through pinned staging buffers owned by the handle,
 * overlapped with the copies and the decode of the neighbouring chunk.  Values outside [0, card_channel) are an
 * error. */
int ibldpc_decode_ib_host_i32(ibldpc_handle h, const int32_t *ch_host, int64_t B, int imax, int early_term,
                              int32_t *out_host, int32_t *i_num_host);

/* Packed host buffers (opt-in; what the BER drivers actually consume is one hard decision per information bit,
 * Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_enc.py:134, discrete_LDPC_decoder_irreg.py:343-349):
 *   ch4_host   (n_var, ceil(B/2)) bytes, frame f of a row in nibble (f & 1) of byte f >> 1 (low nibble = even frame)
 *   bits_host  (rows, ceil(B/8)) bytes, bit (f & 7) of byte f >> 3 = decoded bit (cluster < card_decoder / 2) of the
 *              first `rows` rows (rows = data_len for the irregular decoders, n_var for the regular one)
 * Host->device traffic is half, device->host traffic 1/16 (rows = n_var / 2) of ibldpc_decode_ib_host.
 * Needs card_channel, card_decoder <= 16. */
int ibldpc_decode_ib_host_packed(ibldpc_handle h, const uint8_t *ch4_host, int64_t B, int imax, int early_term,
                                 uint8_t *bits_host, int64_t rows, int32_t *i_num_host);

/* Replaces decode_OpenCL_min_sum (min_sum_decoder_irreg.py:221-287) and
 * decode_OpenCL_belief_propagation (bp_decoder_irreg.py:221-286).  ch_dev / out_dev are
 * (n_var, B) LLR arrays of `dtype` (IBLDPC_F32: fp32 messages, the fast path;
 * IBLDPC_F64: float64 like the reference). */
int ibldpc_decode_llr(ibldpc_handle h, int algo, int dtype, const void *ch_dev, int64_t B, int imax,
                      int early_term, void *out_dev, int32_t *i_num_host, void *stream);

/* Opt-in LAYERED (row-message-passing) schedule of the same two decoders -- SURVEY.md 8(f) rank 4; the reference
 * (min_sum_decoder_irreg.py:242-273, bp_decoder_irreg.py:242-272) only has the flooding schedule, so this entry point has
 * no reference counterpart and its results differ from ibldpc_decode_llr by construction.  One a-posteriori LLR per
 * variable node, updated check by check: X = L - R_old, R_new = checknode(clip150(X)) with the reference's node
 * arithmetic (kernels_min_and_BP.cl:126-167 min-sum, :5-9 box-plus), L = X + R_new; layers = greedy colouring of the
 * checks in index order (checks of a layer share no variable).  At most imax - 1 passes; early_term stops the batch when
 * the syndrome of the hard decisions is zero; out_dev = a-posteriori LLRs; *i_num = passes + 1.  Same buffers as
 * ibldpc_decode_llr.  ibldpc_layer_count returns the number of layers of the handle's code. */
int ibldpc_decode_llr_layered(ibldpc_handle h, int algo, int dtype, const void *ch_dev, int64_t B, int imax,
                              int early_term, void *out_dev, int32_t *i_num_host, void *stream);
int ibldpc_layer_count(ibldpc_handle h, int32_t *n_layers);

/* Replaces return_errors_all_zero (discrete_LDPC_decoder.py:297-300,
 * discrete_LDPC_decoder_irreg.py:343-349) and the host-side comparison of the _enc drivers
 * (Irregular_LDPC_Decoding/WLAN/BER_simulation_OpenCL_enc.py:134): over the first `rows`
 * rows of out_dev (rows_total x B), decoded bit = (out < threshold); counters_host[0] = bit
 * errors, [1] = frames with at least one bit error.  ref_bits_dev == NULL means the all-zero
 * codeword, else uint8 (rows, B) transmitted bits.  Synchronises the stream. */
int ibldpc_count_errors_u8(int device, const uint8_t *out_dev, int64_t rows, int64_t B, int threshold,
                           const uint8_t *ref_bits_dev, int64_t *counters_host, void *stream);

/* Same counters, asynchronous: counters_dev[0] += bit errors, counters_dev[1] += frame errors
 * (int64 device memory), no host synchronisation -- for pipelined BER loops where the per-batch
 * counters are all-reduced on the device and read back one batch late. */
int ibldpc_count_errors_u8_async(int device, const uint8_t *out_dev, int64_t rows, int64_t B, int threshold,
                                 const uint8_t *ref_bits_dev, int64_t *counters_dev, void *stream);

/* LLR twin (min_sum_decoder_irreg.py:290-295, bp_decoder_irreg.py:288-293): bit = (LLR < 0). */
int ibldpc_count_errors_llr(int device, const void *out_dev, int dtype, int64_t rows, int64_t B,
                            const uint8_t *ref_bits_dev, int64_t *counters_host, void *stream);

/* Asynchronous twin of ibldpc_count_errors_llr (counters_dev[0] += bit errors, [1] += frame errors). */
int ibldpc_count_errors_llr_async(int device, const void *out_dev, int dtype, int64_t rows, int64_t B,
                                  const uint8_t *ref_bits_dev, int64_t *counters_dev, void *stream);

/* Replaces the `quantize` kernel (AWGN_Channel_Transmission/kernels_quanti_template.cl:2-27)
 * as launched by quantize_OpenCL (AWGN_Quantizer_BPSK.py:183-199):
 * cluster = #{w in [1,card) : x - limits[w] > 0}, float64 compare.  limits_host: card doubles. */
int ibldpc_quantize(int device, const double *x_dev, int64_t n, const double *limits_host, int card,
                    uint8_t *out_dev, void *stream);

/* Replaces `quantize_LLR` (kernels_quanti_template.cl:29-52): out = llr_host[cluster]. */
int ibldpc_quantize_llr(int device, const double *x_dev, int64_t n, const double *limits_host, int card,
                        const double *llr_host, int dtype, void *out_dev, void *stream);

/* Replaces quantize_direct_OpenCL / quantize_direct_OpenCL_LLR (AWGN_Quantizer_BPSK.py:201-248):
 * u ~ U[0,1) drawn ON THE DEVICE (Philox4x32-10, element i uses counter offset+i of key `seed`)
 * instead of np.random.rand + H2D, then the same inversion-method compare against
 * cdf_host (card = |T|+1 entries).  ibldpc_uniform exposes the very same u for checking. */
int ibldpc_sample_direct(int device, const double *cdf_host, int card, uint64_t seed, uint64_t offset,
                         int64_t n, uint8_t *out_dev, void *stream);
int ibldpc_sample_direct_llr(int device, const double *cdf_host, int card, const double *llr_host,
                             uint64_t seed, uint64_t offset, int64_t n, int dtype, void *out_dev,
                             void *stream);
int ibldpc_uniform(int device, uint64_t seed, uint64_t offset, int64_t n, double *out_dev, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: frames are sharded over one process per GPU; the only collective of the path is the
 * sum of the per-batch counters {bit errors, frame errors, frames, iterations} so that every rank takes the same
 * `while errors < min_errors` decision (Regular_LDPC_Decoding/BPSK/BER_simulation_OpenCL.py:98).
 * NCCL is bound at run time (libnccl.so.2 of the host process, or IBLDPC_NCCL_LIB).
 *   ibldpc_nccl_unique_id     rank 0: 128-byte ncclUniqueId to hand to the other ranks (pipe / file / env)
 *   ibldpc_nccl_init          every rank: join the communicator of `world` ranks on the handle's GPU
 *   ibldpc_allreduce_counters in-place sum of n int64 device counters over all ranks, asynchronous on `stream`
 *   ibldpc_nccl_finalize      destroy the communicator (also done by ibldpc_destroy)
 * --------------------------------------------------------------------------------------------- */
int ibldpc_nccl_unique_id(uint8_t *id128);
int ibldpc_nccl_init(ibldpc_handle h, const uint8_t *id128, int rank, int world);
int ibldpc_allreduce_counters(ibldpc_handle h, int64_t *counters_dev, int n, void *stream);
int ibldpc_nccl_finalize(ibldpc_handle h);

/* Introspection for tests / benchmarks: which[0] = kernel family of the loaded tables (0 = generic path: tables in
 * global memory; 1 = uint8 shared-memory fast path; 2 = packed-nibble fast path; 3 = |T| <= 32 shared-memory family),
 * which[1] = kernels launched by the
 * last decode call, which[2] = persistent grid size, which[3] = dynamic smem bytes. */
int ibldpc_info(ibldpc_handle h, int32_t *which4);

/* Average device time (ms) of the kernels of the last ibldpc_decode_ib call split by phase:
 * ms3[0] = CN kernels, ms3[1] = VN kernels, ms3[2] = iter-0 + output.  Needs
 * ibldpc_set_profiling(h, 1) before the decode call; synchronises. */
int ibldpc_set_profiling(ibldpc_handle h, int on);
int ibldpc_phase_times(ibldpc_handle h, float *ms3, int32_t *launches3);

/* Frames per chunk of the pinned-host pipeline of ibldpc_decode_ib_host; 0 (default) = automatic,
 * equal chunks of at most ~256 MiB of channel values, first and last chunk split 1/4 + 3/4. */
int ibldpc_set_host_chunk(ibldpc_handle h, int frames);

/* ---------------------------------------------------------------------------------------------
 * Transmit side of the BER drivers: systematic encoder, random bits, channel.
 * --------------------------------------------------------------------------------------------- */
#define IBLDPC_ENC_SUBSTITUTION 1 /* H_last (or a row permutation of it) is triangular */
#define IBLDPC_ENC_DENSE 2        /* "Matrix Inverse": dense H_last^-1 from a GF(2) elimination on the host */

typedef struct ibldpc_encoder *ibldpc_encoder_handle;

/* What LDPCEncoder.getLDPCEncoderParamters derives from H (Discrete_LDPC_decoding/LDPC_encoder.py:196-262), in the
 * form the GPU solver consumes.  H = [H_first | H_last], H_last the last M = n_var - n_info columns.  Host pointers.
 *   a_rowptr/a_col        CSR of H_first (M rows, columns in [0, n_info))  -- MatrixA of the reference
 *   substitution          step t solves parity bit var[t] from check equation eq[t]:
 *                           p[var[t]] = s[eq[t]] ^ XOR p[oth[oth_ptr[t] .. oth_ptr[t+1])],   s = H_first x
 *                         (forward / backward substitution and the row-reversed variants of the reference are all
 *                         orderings of this schedule)
 *   dense_inverse         M rows of ceil(M/32) words, bit k of row r = (H_last^-1)[r][k]; p = H_last^-1 s */
typedef struct ibldpc_encoder_desc {
    int32_t n_var, n_info;
    int32_t method; /* IBLDPC_ENC_* */
    const int32_t *a_rowptr, *a_col;
    const int32_t *eq, *var, *oth_ptr, *oth;
    const uint32_t *dense_inverse;
} ibldpc_encoder_desc;

/* Replaces LDPCEncoder.__init__ / setParityCheckMatrix (LDPC_encoder.py:23-38, :192-194): validates and uploads. */
int ibldpc_encoder_create(const ibldpc_encoder_desc *desc, int device, ibldpc_encoder_handle *out);
/* Replaces LDPCEncoder.encode / encode_c (LDPC_encoder.py:86-163) and the per-frame loop of
 * LDPC_BPSK_Transmitter.transmit (LDPC_Transmitter.py:109-121) for a whole batch:
 * bits_dev (n_info, B) uint8 0/1 -> codeword_dev (n_var, B) uint8, first n_info rows = the information bits. */
int ibldpc_encode(ibldpc_encoder_handle h, const uint8_t *bits_dev, int64_t B, uint8_t *codeword_dev, void *stream);
int ibldpc_encoder_destroy(ibldpc_encoder_handle h);
/* Replaces np.random.randint(0, 2, (data_len, msg_at_time)) (LDPC_Transmitter.py:111): Philox4x32-10, counter =
 * offset + element index, bit = top bit of the first output word. */
int ibldpc_random_bits(int device, uint64_t seed, uint64_t offset, int64_t n, uint8_t *out_dev, void *stream);
/* Replaces AWGN_channel.transmission (AWGN_channel.py:32-50, real noise) and, with bits_dev, BPSK_mapping +
 * transmission (LDPC_Transmitter.py:127-132): y = x + sqrt(sigma_n2) n  resp.  y = (1 - 2 bit) + sqrt(sigma_n2) n,
 * n ~ N(0,1) by Box-Muller from Philox4x32-10 (counter = offset + element index).  Exactly one of x_dev / bits_dev. */
int ibldpc_awgn(int device, const double *x_dev, const uint8_t *bits_dev, int64_t n, double sigma_n2, uint64_t seed,
                uint64_t offset, double *y_dev, void *stream);

/* Host-side planning, exposed so that the CPU test-suite can check it without a GPU (no reference counterpart:
 * the reference leaves the OpenCL local size to the runtime, discrete_LDPC_decoder.py:211,219,235-237, and has
 * no host pipeline).
 * ibldpc_plan_geometry: launch geometry of one degree class of the packed-nibble kernels for `resident_ctas`
 *   resident CTA slots of `warps_per_cta` warps, `tiles` frame tiles per row and `n_nodes` nodes:
 *   out3 = {log2(tiles per CTA), tile groups, CTAs per tile group}.
 * ibldpc_host_chunk_schedule: chunk widths (frames) ibldpc_decode_ib_host uses for a batch of B frames of a code
 *   with n_var variable nodes; returns the number of chunks (> capacity: only the first `capacity` are written). */
int ibldpc_plan_geometry(int64_t resident_ctas, int warps_per_cta, int tiles, int n_nodes, int32_t *out3);
int ibldpc_host_chunk_schedule(int64_t B, int64_t n_var, int64_t host_chunk, int early_term, int64_t *widths,
                               int capacity);

const char *ibldpc_last_error(void);
int ibldpc_destroy(ibldpc_handle h);

#ifdef __cplusplus
}
#endif
#endif /* IBLDPC_H */
