/*
 * ldpc_b200.h -- C ABI of libldpc_b200.so: batched normalized-min-sum (NMS) decoding and
 * ordered-statistics decoding (OSD) of short (128,64) LDPC codes on NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference
 * (lgw-frank/Short_LDPC_Decoding_OSD) is Python/TensorFlow and has no FFI of its own; every entry
 * point below names the reference callable it replaces (paths relative to LDPC_128/).  The
 * Python modules in short_ldpc_decoding_osd_b200/ keep those callables' names and signatures and
 * bind this library through ctypes.
 *
 * Conventions
 *  - Plain C types only.  Pointers named *_dev are CUDA device pointers on the handle's device,
 *    pointers named *_host are host pointers.  The caller owns every buffer.
 *  - Every call returns LDPCB_OK (0) or a negative ldpcb_status; nothing throws across the ABI.
 *    ldpcb_last_error() returns a human-readable message for the last failure on that handle.
 *  - Device-pointer calls are asynchronous on `stream` (a cudaStream_t passed as void*, NULL =
 *    default stream).  They do not synchronise, with one exception: calls that need scratch memory
 *    (ldpcb_decode, ldpcb_simulate, ldpcb_select_flagged, ldpcb_osd_pb_decode) keep one scratch buffer per
 *    caller stream; its first use allocates and a later call with a larger batch re-allocates after a
 *    cudaDeviceSynchronize.  Warm a stream up with the largest batch to keep the steady state asynchronous.
 *    *_host calls are synchronous: they return when the results are in the host buffers.
 *  - One handle per (process, device).  A handle is not thread-safe; different handles are
 *    independent, also on different devices in one thread: every entry point makes the handle's
 *    device current for the duration of the call and restores the caller's device on return
 *    (ldpcb_create included).  Calls of ONE handle on different streams may overlap on the device.
 *  - LLR rows are 128 contiguous floats and must be 16-byte aligned (128-bit loads).
 *  - Bit packing is little-endian: bit j of a frame is (w[j >> 5] >> (j & 31)) & 1, with
 *    uint32_t w[4] per frame.  Bit value 1 means "LLR <= 0" (reference: tf.where(x>0,0,1),
 *    Ldpc_128_testing/ms_test.py:39, FS_OSD/convention_osd.py:54).
 *  - There is no CPU fallback: without a CUDA device ldpcb_create() fails.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPCB_ABI_VERSION 1
#define LDPCB_N 128        /* code length the kernels are specialised for   */
#define LDPCB_M 64         /* parity checks                                  */
#define LDPCB_K 64         /* information bits                               */
#define LDPCB_MAX_CHK_DEG 8
#define LDPCB_MAX_VAR_DEG 8
#define LDPCB_MAX_ITERS 64
#define LDPCB_NUM_COUNTERS 16

typedef struct ldpcb_handle ldpcb_t;

typedef enum {
    LDPCB_OK = 0,
    LDPCB_ERR_ARG = -1,      /* NULL pointer, negative size, unknown enum value     */
    LDPCB_ERR_SHAPE = -2,    /* code shape or degree not supported by the kernels   */
    LDPCB_ERR_ALIGN = -3,    /* LLR pointer not 16-byte aligned                     */
    LDPCB_ERR_CUDA = -4,     /* CUDA runtime error (message in ldpcb_last_error)    */
    LDPCB_ERR_CODE = -5,     /* H.G^T != 0 or G not full rank                       */
    LDPCB_ERR_NO_DEVICE = -6 /* no usable CUDA device                               */
} ldpcb_status;

/* TEP (test error pattern) enumeration orders. */
typedef enum {
    /* conventional: weight classes 0..p, inside a class descending sum of MRB indices, ties in
     * lexicographic combination order (PB_OSD/convention_osd.py:13-38).  Index 0 of the MRB is the
     * most reliable position. */
    LDPCB_TEP_CONV = 0,
    /* FS: all-zero pattern first, then per weight lexicographic combinations with the vector
     * reversed, i.e. support {c} -> MRB positions {63-c} (FS_OSD/fs_testing.py:32-49,131-132). */
    LDPCB_TEP_FS = 1
} ldpcb_tep_order;

/* OSD flags (bit mask). */
enum {
    /* Sort ties: default is tf.argsort(|y|, DESCENDING) = stable, lower index first
     * (PB_OSD/pb_testing.py:308-310).  With this flag the order is the exact reverse of
     * tf.argsort(|y|, ASCENDING) (DL_OSD_Testing_serial/ordered_statistics_decoding.py:25-28),
     * i.e. ties put the higher index first.  By matroid duality the most-reliable basis found on G
     * in that order is the complement of the least-reliable basis the reference finds on H. */
    LDPCB_OSD_TIES_HIGH_INDEX_FIRST = 1,
    /* Hard decisions used in the discrepancy come from score_llr instead of order_llr
     * (DL path: ordered_statistics_decoding.py:182-189 scores against the channel LLR while the MRB
     * hard decisions come from the ordering metric).  Default: from order_llr
     * (FS_OSD/convention_osd.py:54-60; PB_OSD/convention_osd.py:54-61). */
    LDPCB_OSD_DISC_HARD_FROM_SCORE = 2,
    /* ldpcb_osd_block_minima only: largest TEP weight in the caller's list (1..4) as a hint, flags |= w << 4; 0 = unknown
     * (treated as 4).  The sweep touches `w` generator rows per TEP instead of 4. */
    LDPCB_OSD_MAXW_SHIFT = 4
};

/* Indices into the uint64 counter block filled by ldpcb_tally / ldpcb_decode_* . */
enum {
    LDPCB_CNT_FRAMES = 0,
    LDPCB_CNT_NMS_FRAME_ERR = 1,   /* NMS hard decision != truth (ms_test.py:40-42,52)            */
    LDPCB_CNT_NMS_BIT_ERR = 2,     /* ms_test.py:41,53                                            */
    LDPCB_CNT_NMS_DETECTED = 3,    /* non-zero syndrome after the last iteration (ms_test.py:51)  */
    LDPCB_CNT_NMS_UNDETECTED = 4,  /* zero syndrome but wrong codeword (ms_test.py:45-50)         */
    LDPCB_CNT_NMS_ITERS = 5,       /* sum of iterations actually run                              */
    LDPCB_CNT_OSD_FRAMES = 6,      /* frames handed to OSD                                        */
    LDPCB_CNT_OSD_FRAME_ERR = 7,   /* OSD codeword != truth (convention_osd.py:67-75)             */
    LDPCB_CNT_OSD_BIT_ERR = 8,
    LDPCB_CNT_FINAL_FRAME_ERR = 9, /* undetected NMS errors + OSD errors                          */
    LDPCB_CNT_FINAL_BIT_ERR = 10,
    LDPCB_CNT_TEPS = 11,           /* TEPs evaluated                                              */
    LDPCB_CNT_PHASE0 = 12          /* 12..15: weight (0..3) of the winning TEP of correct OSD frames
                                      (belonged_phase, convention_osd.py:70-74)                   */
};

int ldpcb_abi_version(void);

/* Number of CUDA devices visible (0 if none / driver missing). Never fails. */
int ldpcb_device_count(void);

/*
 * Build a decoder for one code on one device.  H and G are host arrays of 0/1 bytes, row-major
 * [m*n] and [k*n].  Replaces Code(H_filename) + GL.set_map('code_parameters', code)
 * (Ldpc_128_testing/fill_matrix_info.py:70-129, ldpc_128_testing.py:53-54): builds the edge lists,
 * the packed generator columns and the TEP tables once.  Supported: n=128, m=64, k=64, check degree
 * <= 8, variable degree <= 8, G of full rank with H.G^T = 0.
 */
int ldpcb_create(ldpcb_t** h, const uint8_t* H_host, const uint8_t* G_host, int n, int m, int k, int device);
void ldpcb_destroy(ldpcb_t* h);
/* h may be NULL: returns the message of the last failed ldpcb_create in this thread. */
const char* ldpcb_last_error(ldpcb_t* h);
/* Multiprocessor count of the handle's device (grid sizing information for callers). */
int ldpcb_sm_count(ldpcb_t* h);

/*
 * BPSK/AWGN frame generator.  Replaces testing_data_generating(code, SNR, max_frame)
 * (Testing_data_gen_128/data_generating.py:13-51, AWGN branch, random codewords):
 * sigma = sqrt(1/(2*(k/n)*10^(ebn0_db/10))), message bits uniform, cw = msg.G mod 2,
 * y = (1-2*cw) * (1 + sigma*z), no LLR scaling.  Counter-based: frame f of a run is a pure
 * function of (seed, first_frame + f), so any shard of a run can be generated on any GPU.
 * Philox4x32-10, key = (seed_lo, seed_hi); counter = (frame_lo, frame_hi, block, stream):
 * stream 0 blocks 0..31 give the 128 normals (Box-Muller in fp32, two per pair of words),
 * stream 1 block 0 words 0..1 give the 64 message bits.
 *   llr_dev      [B,128] float, out (may be NULL)
 *   cw_bits_dev  [B,4]   uint32, out (may be NULL): transmitted codeword (the labels)
 */
int ldpcb_gen_frames(ldpcb_t* h, uint64_t seed, uint64_t first_frame, int64_t B, float ebn0_db,
                     float* llr_dev, uint32_t* cw_bits_dev, void* stream);

/*
 * Normalized min-sum BP, flooding schedule.  Replaces Decoder_Layer.call / belief_propagation_op
 * (Ldpc_128_testing/ms_test.py:99-121) with compute_vc (:124-137), compute_cv2 (:180-210),
 * marginalize (:220-228) and the hard decision + syndrome of get_eval (:38-44,51).
 *   alpha_check  softplus(shared_check_weight) (ms_test.py:207-208)
 *   w_vc, w_marg softplus of the bit weights of NMS-2/NMS-3 (ms_test.py:127-131,222-226); 1.0 for NMS-1
 *   early_stop   0 = always `iters` iterations like the reference (ms_test.py:230-232);
 *                1 = stop a frame at the first iteration whose hard decision has zero syndrome
 *   hard_bits_dev   [B,4] uint32 out: hard decision after the last iteration run
 *   iters_used_dev  [B] uint8 out (may be NULL)
 *   syndrome_nz_dev [B] uint8 out (may be NULL): 1 if the final hard decision fails a check
 *   soft_traj_dev   [B,iters+1,128] float out (may be NULL): row 0 = input, row i = posterior after
 *                   iteration i (soft_output_list, ms_test.py:107-110,227); with early_stop the rows
 *                   after the stopping iteration repeat the last posterior
 */
int ldpcb_nms_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int iters, float alpha_check,
                     float w_vc, float w_marg, int early_stop, uint32_t* hard_bits_dev,
                     uint8_t* iters_used_dev, uint8_t* syndrome_nz_dev, float* soft_traj_dev,
                     void* stream);

/*
 * Ordered-statistics decoding of B frames, one warp per frame: reliability sort, GF(2)
 * elimination of the permuted generator matrix, TEP sweep, re-encode, discrepancy, argmin.
 * Replaces swapped_info + identify_mrb + full_gf2elim (PB_OSD/pb_testing.py:231-320,
 * FS_OSD/fs_testing.py:233-322) followed by convention_osd_main (FS_OSD/convention_osd.py:49-77,
 * PB_OSD/convention_osd.py:49-77) for every frame.
 *   order_llr_dev [B,128] ordering metric: sort key |y| and MRB hard decisions
 *   score_llr_dev [B,128] scoring metric (weights |y| of the discrepancy); may alias order_llr_dev
 *   order         0..3 maximum TEP weight (1, 65, 2081, 43745 TEPs)
 *   tep_order     ldpcb_tep_order
 *   flags         LDPCB_OSD_* mask
 *   cw_bits_dev   [B,4] uint32 out: best codeword, ORIGINAL bit positions
 *   best_tep_dev  [B] int32 out (may be NULL): index of the winning TEP in enumeration order (first
 *                 minimum, tf.argmin, convention_osd.py:63)
 *   best_score_q_dev [B] int64 out (may be NULL), score_exp_dev [B] int32 out (may be NULL):
 *                 exact discrepancy of the winner = best_score_q * 2^(score_exp-54).  Scores are
 *                 exact integers: each |y| is scaled by 2^(54-E) (E = frexp exponent of the frame's
 *                 largest |score_llr|) and rounded to nearest-even; see DESIGN.md "exact score".
 *   perm_dev      [B,128] uint8 out (may be NULL): position t of the permuted frame is original
 *                 position perm[t] (pi2 o pi1: 64 MRB positions by descending reliability, then
 *                 64 LRB positions by descending reliability; pb_testing.py:284-302,308-319)
 *   redG_dev      [B,64] uint64 out (may be NULL): row t of reduced_G = [I | P'] is e_t followed by
 *                 the 64 bits of redG[t] (bit l = column 64+l) (pb_testing.py:288-300)
 */
int ldpcb_osd_decode(ldpcb_t* h, const float* order_llr_dev, const float* score_llr_dev, int64_t B,
                     int order, int tep_order, int flags, uint32_t* cw_bits_dev, int32_t* best_tep_dev,
                     int64_t* best_score_q_dev, int32_t* score_exp_dev, uint8_t* perm_dev,
                     uint64_t* redG_dev, void* stream);

/*
 * Same elimination, caller-supplied TEP list cut into blocks; returns per-block minima.
 * Replaces osd.acquire_min over the TEP blocks of a decoding path
 * (DL_OSD_Testing_serial/ordered_statistics_decoding.py:153-162,186-203; blocks from
 * error_pattern_gen :81-98 / nn_testing.py:144-157).
 *   teps_dev      [n_teps] uint32: up to four MRB positions, one per byte, 0xFF = unused; position 0
 *                 is the MOST reliable MRB position (DL index i maps to 63-i)
 *   block_start_dev [n_blocks+1] int32 ascending, block b = teps[block_start[b] .. block_start[b+1])
 *   block_min_q_dev [B,n_blocks] int64 out: minimum exact score of each block
 *   block_arg_dev   [B,n_blocks] int32 out (may be NULL): TEP index of that minimum (first one)
 *   truth_bits_dev  [B,4] uint32 (may be NULL) and truth_score_q_dev [B] int64 out (may be NULL):
 *                 exact discrepancy of the transmitted codeword (discrepancy_sum_truth,
 *                 ordered_statistics_decoding.py:181-184)
 */
int ldpcb_osd_block_minima(ldpcb_t* h, const float* order_llr_dev, const float* score_llr_dev, int64_t B,
                           const uint32_t* teps_dev, int32_t n_teps, const int32_t* block_start_dev,
                           int32_t n_blocks, int flags, int64_t* block_min_q_dev, int32_t* block_arg_dev,
                           int32_t* score_exp_dev, const uint32_t* truth_bits_dev,
                           int64_t* truth_score_q_dev, uint8_t* perm_dev, void* stream);

/*
 * Sweep only, host buffers: the frames are already permuted (64 MRB positions first) and come with the
 * 64 P' words of their systematic generator [I | P'].  Replaces convention_osd_main(wrapped_input)
 * (FS_OSD/convention_osd.py:49-77; PB_OSD/convention_osd.py:49-77 for the 6-tuple form) when the caller
 * brings its own updated_inputs / reduced_G / error_patterns_matrix, e.g. from swapped_info.
 *   upd_order_llr_host [B,128] permuted ordering metric (MRB hard decisions), upd_score_llr_host [B,128]
 *   permuted scoring metric (may alias), redG_host [B,64] uint64, teps_host [n_teps] packed TEPs
 *   cw_bits_host [B,4] out: best codeword in the PERMUTED order of the inputs
 */
int ldpcb_osd_sweep_host(ldpcb_t* h, const float* upd_order_llr_host, const float* upd_score_llr_host,
                         const uint64_t* redG_host, int64_t B, const uint32_t* teps_host, int32_t n_teps, int flags,
                         uint32_t* cw_bits_host, int32_t* best_tep_host, int64_t* best_score_q_host,
                         int32_t* score_exp_host);

/*
 * FS-OSD (fast and scalable OSD) policy on top of the same sort / elimination / sweep machinery.
 * Replaces the per-frame body of fs_osd(snr, beta, selected_ds) (FS_OSD/fs_testing.py:94-165):
 * order-0 acceptance below tau_e (:131-135), order-skip rule with beta (:22-30,137-139), per-TEP
 * tau_e stop -- which does not update the decision, as in the reference (:143-147) -- and tau_psc
 * gated improvement (:148-152), TEPs in generate_sequential_teps order (:32-49).
 *   llr_dev [B,128] channel LLR of the frames (row 0 of the retest record, fs_testing.py:98-99)
 *   order_limit 0..3; tau_e = floor(d_min-1)/2 = 6.5 for d_min 14 (fs_testing.py:92); tau_psc 30;
 *   beta_shift = beta*(n-k) (6.4 for beta 0.1)
 *   cw_bits_dev [B,4] out: optimal_codeword, original bit positions
 *   best_tep_dev [B] out (may be NULL): its index in the FS enumeration
 *   num_teps_dev [B] out (may be NULL): TEPs visited (the reference's num_teps, :130,142)
 *   stop_kind_dev [B] out (may be NULL): 0 order-0 accepted, 1 tau_e stop, 2 skip rule, 3 all orders swept
 */
int ldpcb_osd_fs_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int order_limit, float tau_e, int tau_psc,
                        float beta_shift, uint32_t* cw_bits_dev, int32_t* best_tep_dev, int32_t* num_teps_dev,
                        uint8_t* stop_kind_dev, int64_t* best_score_q_dev, int32_t* score_exp_dev,
                        uint8_t* perm_dev, void* stream);
int ldpcb_osd_fs_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int order_limit, float tau_e, int tau_psc,
                             float beta_shift, uint32_t* cw_bits_host, int32_t* best_tep_host, int32_t* num_teps_host,
                             uint8_t* stop_kind_host);

/*
 * PB-OSD (probability-based OSD) policy.  Replaces the per-frame body of pb_osd(snr, selected_ds)
 * (PB_OSD/pb_testing.py:100-149): best-first TEP order (optimal_tep_sequence, :366-397), promising-probability
 * stop (:128-132, :399-447) and success-probability stop (:137-149, :410-423) with the thresholds of
 * calculate_two_thresholds (:485-500).  order_limit 0..3 (3 is the reference's default, PB_OSD/globalmap.py:42; its
 * 43,745-entry TEP lists live in a global-memory workspace owned by the handle, orders 0..2 keep them in shared memory).
 *   llr_dev [B,128] channel LLR of the frames; snr_db the Eb/N0 the reference passes as `snr` (:50-52)
 *   cw_bits_dev [B,4] out: optimal_codeword, original bit positions
 *   stats_dev [B,4] int32 out (may be NULL): TEPs visited (cost_tep_num or N_max), p_e^pro passes
 *   (counter_suc_sum1), improvements (counter_suc_sum2), list comparisons (memory_sum)
 */
int ldpcb_osd_pb_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int order_limit, float snr_db,
                        uint32_t* cw_bits_dev, int32_t* stats_dev, int64_t* best_score_q_dev,
                        int32_t* score_exp_dev, void* stream);
int ldpcb_osd_pb_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int order_limit, float snr_db,
                             uint32_t* cw_bits_host, int32_t* stats_host);

/* Number of TEPs of an enumeration (1, 65, 2081, 43745 for order 0..3), or a negative status. */
int ldpcb_tep_count(ldpcb_t* h, int order, int tep_order);
/* Copy the enumeration to the host in the packed format described above (n = ldpcb_tep_count). */
int ldpcb_tep_table(ldpcb_t* h, int order, int tep_order, uint32_t* teps_host);

/*
 * Stable compaction of the frames whose flag byte is non-zero (the detected NMS failures that go to
 * OSD: index = tf.where(syndrome != 0), ms_test.py:51; collect_failed_output_selective :55-64).
 *   idx_dev   [B] int32 out: ascending frame indices of flagged frames
 *   count_dev [1] int32 out
 */
int ldpcb_select_flagged(ldpcb_t* h, const uint8_t* flags_dev, int64_t B, int32_t* idx_dev,
                         int32_t* count_dev, void* stream);
/* dst[i, :] = src[idx[i], :] for i < *count_dev (rows of `row_floats` floats, multiple of 4). */
int ldpcb_gather_rows(ldpcb_t* h, const float* src_dev, const int32_t* idx_dev, const int32_t* count_dev,
                      int64_t max_rows, int row_floats, float* dst_dev, void* stream);

/*
 * DIA reliability: out[b,j] = bias + sum_i taps[i] * traj[b,i,j].  The reference's conv_bitwise
 * (DL_OSD_Testing_serial/nn_net.py:174-197) is three bias-free linear Conv1D layers and a Dense(1)
 * per bit, i.e. exactly a (iters+1)-tap FIR + bias; the taps are folded on the host.
 */
int ldpcb_dia_fir(ldpcb_t* h, const float* traj_dev, int64_t B, int n_rows, const float* taps_host,
                  float bias, float* out_dev, void* stream);

/*
 * NMS decode with the DIA reliability fused in: metric_dev[b,j] = bias + sum_{i=0..iters} taps_host[i] * posterior_i[b,j]
 * (posterior_0 = the input LLR), accumulated in registers while the iterations run -- the same arithmetic as
 * ldpcb_nms_decode(soft_traj) followed by ldpcb_dia_fir (one fused multiply-add per row in row order, bias last),
 * without writing the (iters+1) x 128 trajectory to HBM.  Replaces Decoder_Layer.call (ms_test.py:99-101) followed by
 * conv_bitwise.call (DL_OSD_Testing_serial/nn_net.py:190-197) on the frames handed to the DL OSD stage.  Always
 * `iters` iterations (no early stop); taps_host holds iters+1 floats.
 */
int ldpcb_nms_decode_fir(ldpcb_t* h, const float* llr_dev, int64_t B, int iters, float alpha_check, float w_vc,
                         float w_marg, const float* taps_host, float bias, uint32_t* hard_bits_dev,
                         uint8_t* syndrome_nz_dev, float* metric_dev, void* stream);

/*
 * DL sliding-window early termination over the block minima of ldpcb_osd_block_minima.  Replaces the window
 * bookkeeping of osd.sliding_osd (DL_OSD_Testing_serial/ordered_statistics_decoding.py:186-219) with
 * sliding_window_ops (:141-151) and the classifier Predict_outlier_light (nn_net.py:136-149: Dense(w+1, no
 * bias, linear) -> Dense(2, no bias, softmax), stop when p[1] > soft_margin).
 *   block_min_q_dev [B,n_blocks], score_exp_dev [B], truth_score_q_dev [B] (may be NULL => success = 0)
 *   win_width <= 8 (reference: 5), n_blocks <= 128 (reference: decoding_length 30)
 *   W1_host [(w+1)*(w+1)] and W2_host [(w+1)*2] row-major Keras kernels (input index first)
 *   acc_block_size_host [n_blocks+1] cumulative TEP counts (nn_testing.py:156)
 *   success_dev [B] uint8, windows_dev [B], complexity_dev [B] out (each may be NULL);
 *   counters_dev [4] uint64 accumulated (may be NULL): successes, failures, windows_sum, complexity_sum
 */
int ldpcb_dl_window_policy(ldpcb_t* h, const int64_t* block_min_q_dev, const int32_t* score_exp_dev,
                           const int64_t* truth_score_q_dev, int64_t B, int n_blocks, int win_width,
                           const float* W1_host, const float* W2_host, float soft_margin,
                           const int32_t* acc_block_size_host, uint8_t* success_dev, int32_t* windows_dev,
                           int32_t* complexity_dev, uint64_t* counters_dev, void* stream);

/*
 * FER/BER tallies (get_eval, ms_test.py:36-54; convention_osd.py:67-75).  Adds to counters_dev.
 *   nms_bits_dev [B,4], syndrome_nz_dev [B], iters_used_dev [B] (may be NULL): NMS results
 *   final_bits_dev [B,4] (may be NULL): decisions after OSD (equal to nms_bits for frames not sent)
 *   best_tep_dev [B] (may be NULL): winning TEP index for the phase histogram, -1 for frames not sent
 *   truth_bits_dev [B,4]: transmitted codewords
 */
int ldpcb_tally(ldpcb_t* h, const uint32_t* nms_bits_dev, const uint8_t* syndrome_nz_dev,
                const uint8_t* iters_used_dev, const uint32_t* final_bits_dev, const int32_t* best_tep_dev,
                int osd_order, int tep_order, const uint32_t* truth_bits_dev, int64_t B,
                uint64_t* counters_dev, void* stream);

/*
 * Whole hot path on device-resident LLRs: NMS on all frames, then OSD (order `osd_order`, or -1 for
 * none) on the frames with a non-zero syndrome using their channel LLR (row 0 of the 13, as
 * PB_OSD/pb_testing.py:71-72), decisions merged and tallied.  No host synchronisation.
 *   final_bits_dev [B,4] out; syndrome_nz_dev [B] out (may be NULL); best_tep_dev [B] out (may be
 *   NULL, -1 where OSD did not run); truth_bits_dev (may be NULL => no tally); counters_dev
 *   [LDPCB_NUM_COUNTERS] uint64, accumulated (may be NULL)
 */
int ldpcb_decode(ldpcb_t* h, const float* llr_dev, int64_t B, int iters, float alpha_check, float w_vc,
                 float w_marg, int early_stop, int osd_order, int tep_order, uint32_t* final_bits_dev,
                 uint8_t* syndrome_nz_dev, int32_t* best_tep_dev, const uint32_t* truth_bits_dev,
                 uint64_t* counters_dev, void* stream);

/*
 * Monte-Carlo step: generate B frames (ldpcb_gen_frames) and run ldpcb_decode on them without the
 * LLRs ever leaving the GPU.  Replaces the Testing_data_gen_128 -> Ldpc_128_testing -> *_OSD
 * file-coupled chain for FER curves.  counters_dev accumulates.
 */
int ldpcb_simulate(ldpcb_t* h, uint64_t seed, uint64_t first_frame, int64_t B, float ebn0_db, int iters,
                   float alpha_check, float w_vc, float w_marg, int early_stop, int osd_order,
                   int tep_order, uint64_t* counters_dev, void* stream);

/*
 * Host-buffer entry points (what the Python drop-ins call with NumPy arrays): chunked, with the
 * host->device copy of chunk i+1 and the device->host copy of chunk i-1 overlapping the kernels of
 * chunk i on separate streams.  Synchronous.  Pinned host buffers (ldpcb_host_alloc) copy fastest.
 */
int ldpcb_nms_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int iters, float alpha_check,
                          float w_vc, float w_marg, int early_stop, uint32_t* hard_bits_host,
                          uint8_t* iters_used_host, uint8_t* syndrome_nz_host, float* soft_traj_host);
int ldpcb_osd_decode_host(ldpcb_t* h, const float* order_llr_host, const float* score_llr_host, int64_t B,
                          int order, int tep_order, int flags, uint32_t* cw_bits_host,
                          int32_t* best_tep_host, int64_t* best_score_q_host, int32_t* score_exp_host,
                          uint8_t* perm_host, uint64_t* redG_host);
int ldpcb_decode_host(ldpcb_t* h, const float* llr_host, int64_t B, int iters, float alpha_check, float w_vc,
                      float w_marg, int early_stop, int osd_order, int tep_order,
                      uint32_t* final_bits_host, uint8_t* syndrome_nz_host, int32_t* best_tep_host,
                      const uint32_t* truth_bits_host, uint64_t* counters_host);

/*
 * Decoding_model.call in one call (Ldpc_128_testing/ms_test.py:30-34): NMS on every frame with the get_eval tallies
 * (:36-54) and the iters+1 posteriors of the frames left with a non-zero syndrome, ascending frame order -- the
 * 13-rows-per-failure retest records of collect_failed_output_selective (:55-64).  The failures are compacted and
 * re-decoded on the device; the host pays one H2D copy and D2H copies sized by the failure count.
 *   truth_bits_host [B,4] and counters_host [LDPCB_NUM_COUNTERS] (both or neither; counters accumulate)
 *   hard_bits_host [B,4] out, syndrome_nz_host [B] out (may be NULL)
 *   fail_idx_host [max_fail] out: frame indices; fail_traj_host [max_fail, iters+1, 128] out: row 0 = input
 *   n_fail_host out: number of frames with a non-zero syndrome (may exceed max_fail; only the first max_fail are stored)
 */
int ldpcb_nms_retest_host(ldpcb_t* h, const float* llr_host, int64_t B, int iters, float alpha_check, float w_vc,
                          float w_marg, const uint32_t* truth_bits_host, uint32_t* hard_bits_host,
                          uint8_t* syndrome_nz_host, uint64_t* counters_host, int64_t max_fail,
                          int32_t* fail_idx_host, float* fail_traj_host, int64_t* n_fail_host);

/* PCI bus id ("0000:1b:00.0") of a CUDA device, for callers that bind their host threads and pinned buffers to the
 * GPU's NUMA node (/sys/bus/pci/devices/<id>/numa_node) before ldpcb_host_alloc's first touch.  len >= 13. */
int ldpcb_device_pci_bus_id(int device, char* buf, int len);

/* Pinned host memory for the *_host calls. */
int ldpcb_host_alloc(void** p, uint64_t bytes);
int ldpcb_host_free(void* p);

/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
uint64_t ldpcb_launch_count(ldpcb_t* h);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
