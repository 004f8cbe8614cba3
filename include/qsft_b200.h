/* qsft_b200.h -- C ABI of libqsft_b200.so: the B200 (sm_100a) kernels of the q-SFT transform path.
 *
 * The reference (basics-lab/qsft) is pure Python and has NO FFI of its own; the boundary a maintainer would bind is
 * its Python object API (QSFT.transform / SubsampledSignal.subsample / query_args).  Each entry point below replaces
 * one numerical call site of that path; the reference file:line it replaces is cited.  INTEGRATION.md shows the
 * ctypes stub.  Conventions:
 *   - every function returns 0 on success or a negative QSFT_E* code; qsft_last_error() gives a thread-local text.
 *   - all data pointers are DEVICE pointers owned by the caller (e.g. torch tensors' data_ptr()); sizes explicit;
 *     no allocation inside unless stated; `stream` is a cudaStream_t passed as void* (NULL = default stream); calls
 *     are asynchronous on that stream unless stated.
 *   - digit vectors are int8, MSB first (digit 0 is the most significant base-q digit, as qsft/utils.py:74-76).
 *   - complex values are interleaved float (re, im) = complex64.
 *   - decimal indices are unsigned little-endian-in-limbs: limbs==1 -> one uint64; limbs==2 -> {hi, lo} uint64 pair.
 */
#ifndef QSFT_B200_H
#define QSFT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSFT_OK 0
#define QSFT_EINVAL (-1)   /* bad argument (message in qsft_last_error)          */
#define QSFT_ECUDA (-2)    /* CUDA runtime error                                  */
#define QSFT_EUNSUPPORTED (-3)

const char* qsft_last_error(void);
int qsft_version(void);
/* number of kernels launched by this library since load / since the last reset (bench.py "gpu_launches"). */
int64_t qsft_launch_count(void);
void qsft_reset_launch_count(void);

/* Host staging of a digit table for upload (host logic, no device work).  Replaces the int64 -> small-integer handling the
 * reference leaves to NumPy when it keeps the support as locq (n, S) int64 (synt_exp/synt_src/synthetic_signal.py:43-44, 93):
 * `rows` rows of n integer digits (elem_bytes 1 / 2 / 4 / 8; row_stride / col_stride in elements, so either orientation of
 * locq works) -> dst (rows, ld) int8, zero padded, typically page-locked memory that one cudaMemcpyAsync then uploads in the
 * layout qsft_eval_synth* / qsft_peel* expect.  `threads` (1..16) host threads of its own.                                   */
int qsft_host_pack_digits(const void* src, int elem_bytes, int64_t rows, int n, int64_t row_stride, int64_t col_stride,
                          int8_t* dst, int ld, int threads);

/* K1 -- query lattice.  Replaces SubsampledSignal._get_qsft_query_indices (qsft/input_signal_subsampled.py:183-206)
 * + qary_ints (qsft/utils.py:107-108) + qary_vec_to_dec (qsft/utils.py:74-76).
 *   M (n, b) int8 row-major, D (P, n) int8.  For every delay row p and every l in Z_q^b (column index = base-q value
 *   of l, l_0 most significant) writes dec((M l + d_p) mod q):
 *   out_idx  (P, q^b, limbs) uint64   (may be NULL)
 *   out_dig  (P, q^b, ld) int8 digit rows, zero padded to ld >= n, ld % 16 == 0   (may be NULL)           */
int qsft_query_lattice(const int8_t* M, const int8_t* D, int q, int n, int b, int P,
                       uint64_t* out_idx, int limbs, int8_t* out_dig, int ld, void* stream);

/* dec_to_qary_vec (qsft/utils.py:79-84) on device: idx (N, limbs) uint64 -> dig (N, ld) int8, zero padded. */
int qsft_dec_to_qary(const uint64_t* idx, int limbs, int64_t N, int q, int n, int8_t* dig, int ld, void* stream);
/* qary_vec_to_dec (qsft/utils.py:74-76) on device: dig (N, ld) -> idx (N, limbs). */
int qsft_qary_to_dec(const int8_t* dig, int ld, int64_t N, int q, int n, uint64_t* idx, int limbs, void* stream);

/* K2 -- synthetic sparse signal evaluation.  Replaces SyntheticSubsampledSignal.subsample / sampling_function
 * (synt_exp/synt_src/synthetic_signal.py:100-118):  out[m] = sum_s strengths[s] * w^(<qdig[m], loc[s]> mod q).
 *   qdig (N, ld) int8, loc (S, ld) int8 (support digit rows, zero padded), strengths (S) complex64, out (N) complex64.
 *   impl: 0 = best available for the shape, 1 = SIMT (dp4a) kernel, 2 = tcgen05 int8 GEMM + fused epilogue.   */
int qsft_eval_synth(const int8_t* qdig, int64_t N, const int8_t* loc, const float* strengths, int64_t S,
                    int q, int n, int ld, float* out, int impl, void* stream);

/* K1+K2 fused for lattice queries, q = 4 ("lattice-factorised evaluation", SURVEY section 7 option b).  Computes the
 * same samples as qsft_query_lattice + qsft_eval_synth for every delay row of one (M, D) block,
 *     out[p][l] = sum_s a_s w^<k_s, (M l + d_p) mod q>,   l in Z_q^b in lattice order,
 * as a complex matrix product with K = S on the tensor cores: with h_s = M^T k_s and l = (l_hi, l_lo),
 *     out[p][l_hi, l_lo] = sum_s i^<h_hi(s), l_hi> * ( a_s i^(<d_p, k_s> + <h_lo(s), l_lo>) ),
 * left operand exact in int8 (0, +-1), right operand = a_s quantised to 3 balanced base-128 int8 limbs (absolute error
 * <= 2^-21 max|a| per coefficient), int32 accumulation (UTCIMMA kind::i8), limbs recombined in the epilogue.
 *   M (n, b) int8, D (P, n) int8 (unpadded, as for qsft_query_lattice), loc (S, ld) int8, strengths (S) complex64,
 *   out (P, q^b) complex64.  Allocates stream-ordered scratch (cudaMallocAsync): ~ 6 S P q^ceil(b/2) bytes.
 * q = 3 (7 <= b <= 20): the cube roots of unity are integers of Z[w] (w^2 = -1 - w), i.e. 2 x 2 integer matrices with entries
 * 0, +-1 in the basis (1, w): the same factorisation runs as a DENSE int8 GEMM (no 2:4 structure), once for the real and once
 * for the imaginary parts of the strengths, and X = c_1 + c_w w is formed afterwards.
 * qsft_eval_lattice_supported returns 1 when the shape is handled (q == 4: 7 <= b <= 14; q == 3: 7 <= b <= 20;
 * q == 2: 14 <= b <= 28, through the q = 4 kernels with doubled phases; q == 5 / 7: the q = 3 construction with q - 1
 * coordinates, 3 * q^(b - b/2) >= 128 and q^(b - b/2) <= 65535). */
int qsft_eval_lattice_supported(int q, int n, int b, int P, int64_t S);
int qsft_eval_synth_lattice(const int8_t* M, const int8_t* D, const int8_t* loc, const float* strengths, int64_t S,
                            int q, int n, int b, int P, int ld, float* out, void* stream);

/* The same with the precision of the strengths stated.  The exact +-1 / +-i operand meets the strengths as three balanced
 * int8 limbs (20 bits below max|a|: every coefficient comes back with an ABSOLUTE error of ~7e-7 max|a|).  residual_passes
 * = 1 runs a second GEMM pass over the quantisation residual, accumulated onto the first (41 bits; twice the tensor work):
 * needed for 1e-5 RELATIVE accuracy of coefficients much smaller than the largest (reference tolerance, qsft/utils.py:161-164
 * draws |a| in [a_min, a_max]).  0 = never, -1 = decide on the device data (second pass iff min|a| < 0.1 max|a|; reads two
 * floats back, i.e. synchronises the stream once).  qsft_eval_synth_lattice == residual_passes -1.                       */
int qsft_eval_synth_lattice_ex(const int8_t* M, const int8_t* D, const int8_t* loc, const float* strengths, int64_t S,
                               int q, int n, int b, int P, int ld, float* out, int residual_passes, void* stream);

/* Measurement noise on the device: U[e] += sd * (N(0,1) + i N(0,1)) for every complex64 bin e < n_bins, Philox4x32-10
 * keyed by (seed, offset + e / 2): the value added to a bin depends only on (seed, offset, e) -- not on the launch, the
 * device or the rank.  Replaces the host loop of SyntheticSubsampledSignal.get_MDU
 * (synt_exp/synt_src/synthetic_signal.py:120-130, sd = noise_sd / sqrt(2 q^b)) when the reference's NumPy random stream
 * is not required (noise_rng="device"); statistical, not bit, parity with the reference.  U must be 16-byte aligned.   */
int qsft_add_noise(float* U, int64_t n_bins, float sd, uint64_t seed, uint64_t offset, void* stream);

/* K3 -- batched b-dimensional length-q DFT, forward sign, scaled by 1/q^b, in place.  Replaces
 * SubsampledSignal._compute_subtransform + gwht (qsft/input_signal_subsampled.py:264-266, qsft/utils.py:31-36).
 *   x (batch, q^b) complex64.  Index <-> digits MSB first on both sides (C-order reshape [q]*b).            */
int qsft_gwht_batch(float* x, int64_t batch, int q, int b, void* stream);
/* Same transform fused with the multi-GPU all-gather of its output: the final stores also go to the same offsets of
 * n_peers (<= 7) peer buffers (peer_x[r] = device pointer, mapped in this process, to the corresponding rows of rank r's
 * symmetric U buffer), i.e. P2P stores over NVLink instead of a separate collective.  peer_x is a HOST array.        */
int qsft_gwht_batch_bcast(float* x, int64_t batch, int q, int b, float* const* peer_x, int n_peers, void* stream);

/* The same with ONE store per element: mc_x = the address of x in the NVLS MULTICAST mapping of the ranks' symmetric U
 * buffers (torch.distributed._symmetric_memory: handle.multicast_ptr + the offset of x in the buffer).  The last pass stores
 * with multimem.st; NVSwitch replicates every store to all ranks of the mapping, this one included, so each byte leaves the
 * GPU once instead of once per peer.  Callers fall back to qsft_gwht_batch_bcast where no multicast mapping exists.       */
int qsft_gwht_batch_mcast(float* x, int64_t batch, int q, int b, float* mc_x, void* stream);

/* ... and the scatter form for the bin-sharded peel (qsft_peel_blocks_sharded): rank r only reads the bins
 * [r * per_bins, (r + 1) * per_bins) of every row, so the last pass stores each element to the ONE rank that owns its bin --
 * rank_x[r] = the address of these rows in rank r's symmetric buffer (rank_x[rank] = x) -- an all-to-all with 1 / world of
 * the all-gather's NVLink traffic.  After it this rank's buffer holds ITS bins of every rank's rows (other bins of other
 * ranks' rows are not written).  per_bins must be the peel's: ceil(q^b / world / 128) * 128.                               */
int qsft_gwht_batch_scatter(float* x, int64_t batch, int q, int b, float* const* rank_x, int world, int rank, int64_t per_bins,
                            void* stream);

/* Verification helper (host only, no device work): work item of ticket number `ticket` in the single-launch two-pass
 * q = 4 transform (k3_q4_twopass_kernel): block, tile inside its pass, and whether it belongs to the strided (second)
 * pass.  A strided tile of block k waits for all tiles1 contiguous tiles of block k; tests check that those always
 * carry lower tickets and that the order is a bijection, for every lag.                                          */
int qsft_k3_ticket_decode(uint32_t ticket, int64_t nblocks, int tiles1, int tiles2, int lag, int64_t* blk, int* tile,
                          int* strided);

/* K4 -- peeling decoder.  Replaces the loop of QSFT.transform (qsft/qsft.py:151-241) and the singleton detectors
 * (qsft/reconstruct.py:12-31, 100-113, 34-51 + qsft/ReedSolomon.py:26-48).
 *
 * Problem description shared by the peel entry points. */
typedef struct {
    int q, n, b;          /* alphabet, signal dimension, subsampling dimension (B = q^b bins per group)            */
    int C, P, P_src;      /* groups, delay rows per group (P = R * P_src), rows of the source delay matrix         */
    int channel;          /* 0 = identity (noiseless angles, reconstruct.py:12-31), 1 = nso1 (reconstruct.py:100-113),
                             2 = nso2 (hard decision, reconstruct.py:116-129 + angle_q utils.py:104-105; the reference's
                             QSFT.transform hard-codes nso1 at qsft.py:171, so this is only reachable by asking for it)  */
    int source;           /* 0 = identity, 1 = coded (Reed-Solomon syndrome decode, ReedSolomon.py:26-48)          */
    int rs_t, rs_s;       /* coded only: error capability t, extension degree s (P_src = 2*t*s + 1)                */
    int ld;               /* row stride (bytes) of MT, D and find_k digit rows: >= n, multiple of 16, zero padded  */
    float cutoff;         /* per-delay energy threshold, qsft.py:124-126                                           */
    const int8_t* MT;     /* (C, b, ld)  row i of group c = column i of M_c                                        */
    const int8_t* D;      /* (C, P, ld)                                                                            */
    const int32_t* rs_exp;/* coded only: GF(q^s) antilog table, 2*(q^s-1) entries (device)                         */
    const int32_t* rs_log;/* coded only: GF(q^s) log table, q^s entries (device)                                   */
} qsft_peel_desc;

/* One classification pass over bins [j_begin, j_end) of every group (qsft.py:162-188).
 *   U (C, P, B) complex64.  For every singleton appends a find f = counters[0]++ (atomic, not zeroed here):
 *   find_cj[f] = c * B + j, find_k (f, ld) digits, find_rho (f) complex64, find_round[f] = round (may be NULL);
 *   find_id (C, B) int32 gets the find number, or -1 for zerotons / multitons.
 *   counters[1] += number of multitons.  Finds beyond max_finds are counted but not stored (caller must check). */
int qsft_peel_classify(const qsft_peel_desc* d, const float* U, int64_t j_begin, int64_t j_end,
                       int64_t* find_cj, int8_t* find_k, float* find_rho, int32_t* find_round, int32_t* find_id,
                       int64_t max_finds, int round, unsigned long long* counters, void* stream);

/* Peel finds [f_begin, f_begin + n_finds) off U (qsft.py:209-241).  With dedupe != 0 a find is applied only if it
 * is the LAST find of its k in (c, j) order (ball_values "last wins", qsft.py:215), checked through find_id /
 * find_k of the higher groups; with dedupe == 0 every listed find is applied.  Only bins in [j_begin, j_end) are
 * touched (bin-sharded multi-GPU peeling).  owner_count (may be NULL) += number of finds applied.               */
int qsft_peel_apply(const qsft_peel_desc* d, float* U, int64_t j_begin, int64_t j_end,
                    const int64_t* find_cj, const int8_t* find_k, const float* find_rho,
                    const int32_t* find_id, int64_t f_begin, int64_t n_finds, int dedupe,
                    unsigned long long* owner_count, void* stream);

/* Distinct-k output of the peel (qsft.py:247-255 averages every rho recorded for the same k).  All device memory. */
typedef struct {
    int32_t* seen0;      /* (B) workspace: chain heads keyed by k's bin in group 0; must be ZERO before round 1        */
    int8_t* uniq_k;      /* (max_uniq, ld) digits of each distinct k                                                   */
    float* uniq_sum;     /* (max_uniq) complex64: sum of rho over all finds of that k (mean = sum / count)             */
    int32_t* uniq_cnt;   /* (max_uniq) number of finds of that k                                                       */
    int64_t* uniq_key;   /* (max_uniq) (round << 48) | (c * B + j) of the first find = the reference's first-seen order */
    int32_t* uniq_next;  /* (max_uniq) workspace (chain links)                                                         */
    int64_t max_uniq;
} qsft_uniq;

/* The distinct-k list in the result layout of QSFT.transform (qsft/qsft.py:247-255: first-seen order, mean over the finds of
 * a k): entry i of the outputs is list entry order[i] (`order` = the entries sorted by uniq_key, device pointer, n_uniq int64)
 * -- k_out (n_uniq, n) int8 without padding, mean_out (n_uniq) complex128 = uniq_sum / uniq_cnt, cnt_out (n_uniq) int32.
 * One kernel in place of the gathers, the slice copy and the host-side division.                                          */
int qsft_peel_distinct(const qsft_uniq* uq, const int64_t* order, int64_t n_uniq, int n, int ld, int8_t* k_out,
                       double* mean_out, int32_t* cnt_out, void* stream);

/* Collapse finds [f_begin, f_begin + n_finds) of round `round` (find_id must be the table of that round) into the
 * distinct-k list; counters[4] is the running number of distinct k (atomic).                                     */
int qsft_peel_reduce(const qsft_peel_desc* d, const int64_t* find_cj, const int8_t* find_k, const float* find_rho,
                     const int32_t* find_id, int64_t f_begin, int64_t n_finds, int round, const qsft_uniq* uq,
                     unsigned long long* counters, void* stream);

/* Whole single-GPU peel loop (qsft.py:151-241): classify / apply rounds until the reference's stop rule
 * (no multitons or no singletons, or 15 rounds, or q^n peels).  SYNCHRONOUS (reads round counters back).
 *   Outputs: finds grouped by round (order inside a round is unspecified): find_cj / find_k / find_rho /
 *   find_round as above.  *n_finds_out = total finds, *n_rounds_out = rounds; with uq != NULL also the distinct-k list
 *   (*n_uniq_out entries).  Workspaces: find_id (C, B) int32, counters (>= 8 x u64, device).
 *   The on-device loop hands find slots out to its warps in chunks: *n_finds_out counts SLOTS, and a slot f that was
 *   not used carries find_cj[f] = -1 (skip it); the distinct-k list has no such gaps.                                                                            */
int qsft_peel(const qsft_peel_desc* d, float* U, int64_t* find_cj, int8_t* find_k, float* find_rho,
              int32_t* find_round, int32_t* find_id, int64_t max_finds, unsigned long long* counters,
              const qsft_uniq* uq /* may be NULL */, int64_t* n_finds_out, int64_t* n_uniq_out, int* n_rounds_out,
              void* stream);

/* The same loop on bins that are NOT one contiguous (C, P, B) array: blocks[c * R + r] (HOST array of C * R DEVICE
 * pointers, R = P / P_src) -> the (P_src, ldU) complex64 rows of group c, repeat r, bin index contiguous, row stride ldU
 * elements -- what SubsampledSignal.get_MDU hands to QSFT.transform (qsft/input_signal_subsampled.py:225-262: per-block
 * row lists in a random group / repeat order; the reference vstacks copies of them, qsft.py:115-121).  The bins are only
 * read: no private copy, no vstack.  Runs the persistent on-device loop (one cooperative kernel, every round on the
 * device); returns QSFT_EUNSUPPORTED (-3, nothing done) when the shape does not fit it -- C * R > 16, P_src > 256 or a
 * tile of 16 bins x P rows beyond the shared memory -- and the caller then assembles U and calls qsft_peel.
 *   ASYNCHRONOUS form: n_finds_out == n_rounds_out == NULL -- the loop is queued on `stream` and the call returns at once
 *   (nothing is read back, so a stream of transforms keeps the GPU busy while the host prepares the next one); the caller
 *   reads `counters` (device) when it needs the outcome: [4] distinct k, [5] rounds, [6] != 0: find buffer too small,
 *   [7] find slots used.                                                                                             */
int qsft_peel_blocks(const qsft_peel_desc* d, const float* const* blocks, int64_t ldU, int64_t* find_cj, int8_t* find_k,
                     float* find_rho, int32_t* find_round, int32_t* find_id, int64_t max_finds, unsigned long long* counters,
                     const qsft_uniq* uq /* may be NULL */, int64_t* n_finds_out, int64_t* n_uniq_out, int* n_rounds_out,
                     void* stream);

/* The on-device loop sharded over the GPUs of one box (one process per GPU, every rank holds all of U): rank `rank` of
 * `world` classifies the bins [rank, rank + 1) * ceil(B / world / 128) * 128 of every group; after every round's
 * classification the kernel pushes the rank's finds into the peers' copies of the find list over NVLink, waits for theirs
 * and links the round's finds of ALL ranks -- the per-round exchange of (k, rho) lists (qsft.py:209-241 on shards) happens
 * inside the one persistent kernel, with no host round trip and no NCCL call.  Every rank ends with the complete
 * distinct-k list (uq).
 *   ws: this rank's workspace of qsft_peel_sharded_workspace_bytes(...) bytes in SYMMETRIC (peer-mapped) memory, identical
 *   layout on every rank, `peers[p]` = address of rank p's workspace as mapped into this process (peers[rank] = ws).  The
 *   first 8 KB (control block) must be zero before the first call; `epoch` must be the same on all ranks and larger at
 *   every call on the same workspace.  All ranks of the call must have finished the previous call on the workspace (a
 *   barrier between peels; in the transform path K3's barrier is one).  Returns QSFT_EUNSUPPORTED like qsft_peel_blocks;
 *   n_uniq_out == n_rounds_out == NULL selects the asynchronous form described there.                                  */
typedef struct {
    int rank, world;
    void* const* peers;     /* [world] */
    uint32_t epoch;
} qsft_shard;
int64_t qsft_peel_sharded_workspace_bytes(const qsft_peel_desc* d, int64_t max_finds);
int qsft_peel_blocks_sharded(const qsft_peel_desc* d, const float* const* blocks, int64_t ldU, const qsft_shard* shard,
                             int64_t max_finds /* slots over all ranks */, unsigned long long* counters,
                             const qsft_uniq* uq, int64_t* n_uniq_out, int* n_rounds_out, void* stream);

/* The reference's public detector entry point for a batch of columns.  Replaces reconstruct.singleton_detection
 * (qsft/reconstruct.py:132-168): channel stage noiseless (:12-31) / nso1 (:100-113) / nso2 (:116-129), then the
 * optional coded source stage (:34-51 + qsft/ReedSolomon.py:26-48).
 *   cols (N, P) complex64, one column of U per ROW (element p of column c at cols[c * P + p]), P = R * P_src.
 *   k_out (N, ld_out) int8: source 0 -> the P_src - 1 detected symbols, source 1 -> the n decoded digits (all zero on
 *   decoder failure, like galois' unchanged zero codeword); zero padded to ld_out.  n is only read for source 1.   */
int qsft_singleton_detect(const float* cols, int64_t N, int q, int n, int P, int P_src, int channel, int source,
                          int rs_t, int rs_s, const int32_t* rs_exp, const int32_t* rs_log, int8_t* k_out, int ld_out,
                          void* stream);

/* Replaces reconstruct.singleton_detection_mle (qsft/reconstruct.py:54-84) for N columns sharing one candidate set:
 *   cols (N, P) complex64 as above, S (P, K) complex64 row-major = S_slice (signature of candidate k in column k).
 *   k_sel[c] = argmin_k || col_c - (<S_k, col_c> / P) S_k ||_2 (first minimum), residual[c] = that norm (may be NULL).
 * The caller maps k_sel through its `selection` list.  (QSFT.transform in the reference never reaches this detector:
 * it does not pass selection / S_slice; it is a public function of reconstruct.py and is kept callable.)             */
int qsft_detect_mle(const float* cols, int64_t N, int P, const float* S, int K, int32_t* k_sel, float* residual,
                    void* stream);

/* Closed-form bins (verification helper): U[p][j] = sum_{s: M^T k_s = j} a_s w^{<d_p,k_s>}.
 * Used by tests and by the peel benchmark to fill U without sampling.  U (P, B) must be zeroed by the caller. */
int qsft_closed_form_bins(const int8_t* MT /* (b, ld) */, const int8_t* D /* (P, ld) */, int q, int n, int b, int P,
                          const int8_t* loc, int ld, const float* strengths, int64_t S, float* U, void* stream);

#ifdef __cplusplus
}
#endif
#endif
