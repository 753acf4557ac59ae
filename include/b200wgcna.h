/* b200wgcna -- C ABI of the B200-native (sm_100a) network-construction hot path of PyWGCNA.
 *
 * This header is the drop-in boundary.  The reference (mortazavilab/PyWGCNA 2.2.0) has no
 * FFI: its "operator interface" for this path is three static methods that findModules looks
 * up on the class by name (PyWGCNA/wgcna.py:278, :310, :320).  Each entry point below names
 * the reference lines whose arithmetic it replaces; pywgcna_b200/wgcna.py binds them with
 * ctypes behind the reference's own Python signatures.
 *
 * Conventions
 *   - plain pointers and sizes only; all matrices are IEEE float64, C order (row-major);
 *   - every function returns 0 on success or a negative B200W_E_* code; the message of the
 *     last failure on the calling thread is b200w_last_error();
 *   - "host" entry points take HOST pointers and do H2D -> kernels -> D2H on `device`;
 *     "dev" entry points take DEVICE pointers valid on `device`, enqueue on `stream`
 *     (a cudaStream_t passed as void*, NULL = default stream) and do not synchronise;
 *   - there is no CPU fallback: without a usable sm_100 device every compute call fails
 *     with B200W_E_NO_DEVICE.
 *   - row blocks: [row0, row0 + nrows) selects the output rows one GPU owns (multi-GPU
 *     row-block sharding); nrows == n, row0 == 0 is the whole matrix.
 */
#ifndef B200WGCNA_H
#define B200WGCNA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200W_VERSION 100

/* error codes */
#define B200W_OK 0
#define B200W_E_INVALID -1   /* bad argument (enum value, size, NULL pointer) */
#define B200W_E_NO_DEVICE -2 /* no CUDA device / not an sm_100 part */
#define B200W_E_CUDA -3      /* a CUDA runtime call or kernel failed */
#define B200W_E_NOMEM -4     /* device allocation failed */
#define B200W_E_UNSUPPORTED -5

/* networkTypes / adjacencyTypes, index order of wgcna.py:53-54 */
#define B200W_NET_UNSIGNED 0
#define B200W_NET_SIGNED 1
#define B200W_NET_SIGNED_HYBRID 2
/* TOMTypes wgcna.py:55, TOMDenoms wgcna.py:56 */
#define B200W_TOM_UNSIGNED 0
#define B200W_TOM_SIGNED 1
#define B200W_DENOM_MIN 0
#define B200W_DENOM_MEAN 1
/* what the TOM epilogue writes */
#define B200W_OUT_TOM 0      /* TOM, diagonal 1                 (wgcna.py:1153-1156) */
#define B200W_OUT_DISS 1     /* 1 - TOM                         (wgcna.py:323)       */
#define B200W_OUT_DISS_R8 2  /* rint((1 - TOM) * 1e8) / 1e8     (wgcna.py:323-324)   */
/* condensed outputs (whole matrix only): the n (n - 1) / 2 entries above the diagonal, row by row --
 * exactly what scipy's squareform(dissTOM, checks=False) hands to linkage (wgcna.py:326-328) */
#define B200W_OUT_DISS_COND 3     /* 1 - TOM, condensed                */
#define B200W_OUT_DISS_R8_COND 4  /* round(1 - TOM, 8), condensed      */
/* arithmetic of the A.A contraction */
#define B200W_PREC_AUTO 0    /* = B200W_PREC_I8 */
#define B200W_PREC_F64 1     /* FP64 tensor-core DMMA (mma.sync f64)                    */
#define B200W_PREC_I8 2      /* tcgen05 kind::i8, exact int32 accumulation of fixed-point
                                digit slices (row-scaled, 5 slices = 39 bits), the 15 digit-pair
                                products of weight >= 4: max-abs error on TOM ~5e-12      */
#define B200W_PREC_I8_FAST 3 /* the same planes, only the 10 digit-pair products of weight >= 5:
                                2/3 of the tensor-core work, max-abs error on TOM ~1e-9 (the
                                north-star tolerance is 1e-6); single-GPU entry points only,
                                the sharded contraction always runs B200W_PREC_I8          */

int b200w_version(void);
const char* b200w_last_error(void);
/* number of visible CUDA devices with compute capability 10.x (0 if none) */
int b200w_device_count(void);

/* ------------------------------------------------------------------ host entry points */

/* Connectivity of the soft-threshold sweep: replaces the block loop of
 * WGCNA.pickSoftThreshold, wgcna.py:873-920 (np.corrcoef :882, transforms :884-889,
 * diag<-1 :897, power chain and nansum :900-916).
 *   X        m x n expression (rows samples, columns genes)
 *   steps    P power STEPS: steps[j] = powers[j] - powers[j-1] for the ascending power
 *            vector (wgcna.py:901-903)
 *   k_out    P x n : k_out[j*n + g] = datk[g, j] of the reference (transposed layout)
 *   bad_out  optional (may be NULL) n flags: 1 for genes whose np.corrcoef row is NaN
 *            (a non-finite sample or zero variance) */
int b200w_connectivity_sweep(const double* X, int64_t m, int64_t n, const double* steps, int P,
                             int net_type, double* k_out, uint8_t* bad_out, int device);

/* The sweep plus, still on the device, everything pickSoftThreshold derives from k[:, power] except
 * the two 10-point regressions (wgcna.py:924-932, :1017-1034): per power a row of
 *   stats[0] min, [1] max, [2] median (statistics.median), [3] float64 sum (informational),
 *   [4 .. 4+nb) bin counts and [4+nb .. 4+2nb) bin sums of pd.cut(k, nb)  (right-closed bins, first edge
 *   lowered by 0.1 % of the range, numpy linspace arithmetic),
 * and buckets[p][e][2] = exact int64 sums of the low / high 27-bit halves of the 53-bit mantissas of all
 * k with frexp exponent e - 1100 (b200w_sft_exp_buckets() entries per power), from which the host forms
 * the exactly rounded statistics.mean (wgcna.py:930).  nbreaks <= 16.  k is not copied to the host. */
int b200w_sft_exp_buckets(void);
int b200w_soft_threshold_stats(const double* X, int64_t m, int64_t n, const double* steps, int P,
                               int net_type, int nbreaks, double* stats_out, int64_t* buckets_out,
                               int device);

/* GENE-MAJOR input: the same two calls for an expression matrix laid out genes x samples (Xt is n x m, C order) --
 * what an F-ordered m x n array is, e.g. DataFrame.to_numpy() when pandas keeps its block gene-major.  Accepting it
 * saves the caller a transposition of the whole matrix; the results are bit-identical to the sample-major calls. */
int b200w_soft_threshold_stats_gm(const double* Xt, int64_t m, int64_t n, const double* steps, int P, int net_type,
                                  int nbreaks, double* stats_out, int64_t* buckets_out, int device);
int b200w_adjacency_gm(const double* Xt, int64_t m, int64_t n, int net_type, double power, int64_t row0,
                       int64_t nrows, double* A_out, int device);

/* WGCNA.adjacency, wgcna.py:1097-1111 (corrcoef, transform, ** power).
 *   A_out    nrows x n row block of the n x n adjacency */
int b200w_adjacency(const double* X, int64_t m, int64_t n, int net_type, double power,
                    int64_t row0, int64_t nrows, double* A_out, int device);

/* The three reductions WGCNA.checkAdjMat needs, wgcna.py:1135-1138, in one pass over A.  stats holds SIX doubles:
 *   stats[0] = max |A - A^T| over non-NaN pairs, stats[1] = min, stats[2] = max (NaN ignored),
 *   stats[3] = number of NaN entries (numpy's max/min return NaN if any is present, which
 *   makes every check of the reference pass; the host logic reproduces that),
 *   stats[4] = max |A - A^T| after nan_to_num (what TomSimilarityFromAdj is handed by TOMsimilarity, :1185),
 *   stats[5] = number of (i, j) whose NaN-ness differs from (j, i). */
int b200w_check_adjacency(const double* A, int64_t n, double* stats, int device);

/* WGCNA.TOMsimilarity / TomSimilarityFromAdj, wgcna.py:1185 (NaN -> 0), :1143 (diag <- 0),
 * :1145 (A @ A), :1146-1151 (k, min/mean denominator), :1153/:1155, :1156 (diag <- 1).
 * The input is not modified.  out is the nrows x n row block.
 * The contraction is L = A @ A with k_i = row sums and k_j = COLUMN sums (:1145-1147): a matrix that is
 * symmetric to 1e-12 (everything checkAdjMat lets through without a NaN) runs the mirrored schedule on one
 * operand, any other takes its column operand from A^T.
 * check == 0 is TomSimilarityFromAdj called directly: no nan_to_num, so a NaN makes every entry of its row
 * and of its column NaN (diagonal 1), exactly as numpy's A @ A and plain sums do.
 * With check != 0 (TOMsimilarity) the checkAdjMat reductions decide first (stats4 = the first four values
 * of b200w_check_adjacency, may be NULL) and the call returns, without computing,
 *   B200W_CHECK_ASYMMETRIC  if max|A - A^T| > 1e-12                     (wgcna.py:1135)
 *   B200W_CHECK_RANGE       if min(A) < lo or max(A) > 1, lo = -1 for the signed TOM else 0 (:1137)
 * -- unless A holds a NaN, which makes both numpy comparisons of the reference False. */
#define B200W_CHECK_ASYMMETRIC 1
#define B200W_CHECK_RANGE 2
int b200w_tom(const double* A, int64_t n, int tom_type, int tom_denom, int out_mode, int precision,
              int64_t row0, int64_t nrows, double* out, int check, double* stats4, int device);

/* Fused chain adjacency -> TOM with the adjacency kept in HBM:
 * equivalent to b200w_adjacency followed by b200w_tom on its result.  A_out may be NULL. */
int b200w_network(const double* X, int64_t m, int64_t n, int net_type, double power, int tom_type,
                  int tom_denom, int out_mode, int precision, double* A_out, double* tom_out,
                  int device);

/* Pearson correlation of every column of X (m x n) with every column of Y (m x M): the n x M block
 * CalculateSignedKME takes out of np.corrcoef(datExpr.T, datME.T) (wgcna.py:3486-3489).  C_out n x M. */
int b200w_cross_correlation(const double* X, int64_t m, int64_t n, const double* Y, int64_t M,
                            double* C_out, int device);
int b200w_dev_cross_correlation(const double* dZx, const uint8_t* d_badx, int64_t n, const double* dZy,
                                const uint8_t* d_bady, int64_t M, int64_t m, double* dC, int device,
                                void* stream);

/* Average-linkage clustering of the genes, the call that follows the TOM in findModules:
 * scipy.cluster.hierarchy.linkage(squareform(dissTOM), method="average") (wgcna.py:326-328), with scipy's exact
 * result -- nearest-neighbour-chain merge order including its tie rules, Lance-Williams heights, stable sort by
 * height, union-find labels (scipy/cluster/_hierarchy.pyx, scipy==1.14.0 pinned by the reference).
 *   y_condensed  n (n - 1) / 2 distances in squareform order;  Z_out  (n - 1) x 4 linkage matrix.
 * b200w_dev_linkage_average works on the n x n square form in HBM (symmetric, destroyed) and leaves the merges
 * in chain order (x, y, height, size) in d_Zraw; b200w_linkage_finish (host) sorts and labels them in place. */
int b200w_linkage_average(const double* y_condensed, int64_t n, double* Z_out, int device);
int b200w_dev_linkage_average(double* d_D, int64_t n, double* d_Zraw, int device, void* stream);
int b200w_dev_condensed_to_square(const double* d_y, int64_t n, double* d_D, int device, void* stream);
int b200w_linkage_finish(double* Z, int64_t n);

/* WGCNA.intramodularConnectivity, wgcna.py:3431-3445: diag <- 0, NaN -> 0, per-module
 * within-module row sums.  labels[n] are module ids in [0, nmod); small[nmod] != 0 marks
 * modules with < 3 genes (kWithin = 0, :3436).  out: kTotal[n], kWithin[n]. */
int b200w_intramodular(const double* A, int64_t n, const int32_t* labels, int nmod,
                       const uint8_t* small, double* k_total, double* k_within, int device);

/* Page-locked host buffers from a recycling pool, for the n x n results: a result written into
 * such a buffer leaves the GPU by DMA at PCIe speed and, handed back as the input of the next call
 * (adjacency -> TOMsimilarity), returns the same way.  Caller buffers that are NOT page-locked are
 * accepted everywhere (they are page-locked for the duration of the call, or staged).
 * b200w_host_free returns the buffer to the pool; b200w_host_trim releases the cached ones. */
void* b200w_host_alloc(int64_t bytes);
void b200w_host_free(void* p);
void b200w_host_trim(void);
/* The host entry points keep their multi-GB device workspaces (adjacency, operand, result) cached
 * between calls instead of cudaMalloc / cudaFree each time; this releases the cached ones. */
void b200w_dev_trim(void);

/* Arithmetic of the correlation product Z Z^T behind the sweep and the adjacency (np.corrcoef at wgcna.py:882,
 * :1097), process wide:
 *   B200W_PREC_I8  (default, = B200W_PREC_AUTO; env B200WGCNA_CORR_PRECISION=i8): tcgen05 kind::i8 MMAs on six
 *                  int8 digit planes of the standardised genes (46-bit fixed point per gene, exact int32
 *                  accumulation, TMA fed; corr_i8.cu) -- max-abs error against np.corrcoef ~2e-13;
 *   B200W_PREC_F64 (env ...=f64): FP64 DMMA (mma.sync), ~1e-15.
 * Step lists other than 1 / 2 between consecutive powers always use the FP64 engine. */
int b200w_set_correlation_precision(int precision);

/* ---------------------------------------------------------------- device entry points */

/* bytes of scratch the dev_* calls need are allocated stream-ordered internally. */

/* centre and scale every gene to unit norm: Z[g, s] = (X[s, g] - mean_g) / ||X[:, g] - mean_g||,
 * Z is n x ldk (ldk = b200w_padded_k(m), zero padded); bad[g] = 1 for genes whose
 * correlation row is NaN in np.corrcoef (non-finite sample or zero variance). */
int64_t b200w_padded_k(int64_t m);
int b200w_dev_standardize(const double* dX, int64_t m, int64_t n, double* dZ, uint8_t* d_bad,
                          int device, void* stream);
/* the same for gene-major input (dXt is n x m) */
int b200w_dev_standardize_gm(const double* dXt, int64_t m, int64_t n, double* dZ, uint8_t* d_bad,
                             int device, void* stream);
int b200w_dev_sweep(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                    const double* h_steps, int P, int net_type, int64_t col0, int64_t ncols,
                    double* d_k /* P x ncols */, int device, void* stream);
/* The sweep that also leaves the transformed similarity behind: d_sim = rows [col0, col0 + ncols) of the
 * symmetric n x n matrix |cor| / (1 + cor) / 2 / max(cor, 0) as WGCNA.adjacency forms it before `** power`
 * (PyWGCNA/wgcna.py:1097-1107; NaN rows and columns kept, diagonal not forced), ncols x n row-major.
 * The reference computes np.corrcoef twice, in pickSoftThreshold (:882) and again in adjacency (:1097);
 * with this pair the second product becomes b200w_dev_adjacency_from_similarity, an in-place elementwise
 * pass (A = S ** power, :1111) whose result is bitwise equal to b200w_dev_adjacency's for the same rows. */
int b200w_dev_sweep_similarity(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                               const double* h_steps, int P, int net_type, int64_t col0, int64_t ncols,
                               double* d_k /* P x ncols */, double* d_sim /* ncols x n */, int device,
                               void* stream);
int b200w_dev_adjacency_from_similarity(double* d_sim_inout /* nrows x n */, int64_t nrows, int64_t n,
                                        double power, int device, void* stream);
/* The sweep for several GPUs with the SYMMETRIC schedule (tcgen05 engine, power steps 1 / 2, P <= 20), fused with the
 * exchange of the similarity.  Rank r owns the genes [r * per, (r + 1) * per) (per % 256 == 0).  The unordered pairs of
 * 128-gene blocks are dealt out over the ranks; the rank that gets a pair forms its correlation tile once and uses it for
 * both gene sets: column sums for its own genes, row sums for the other block's genes, and the transformed similarity
 * stored transposed into its own rows and directly into the row owner's rows -- peer_sim[q] is rank q's per x n
 * similarity / adjacency block (peer-mapped for q != rank, b200w_ipc_open): half the MMAs and power chains of
 * b200w_dev_sweep_similarity on a column block.  d_kpart (P x n) receives this rank's PARTIAL connectivities of all
 * genes; the caller sums them over the ranks (all-reduce) -- which is also the point after which every rank's
 * similarity block is complete. */
int b200w_dev_sweep_dist(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, const double* h_steps, int P,
                         int net_type, int world, int rank, int64_t per, void* const* peer_sim /* host array */,
                         double* d_kpart, int device, void* stream);
/* Host only (no device): the tile list of b200w_dev_sweep_dist for one rank as (128-row block, 64-column tile of the
 * rank's column block) pairs, in launch order.  Returns the number of tiles (-1: bad arguments) and writes them to
 * out_xy[2 * count] when capacity suffices.  For tests of the schedule: over the ranks every pair of blocks exactly once. */
int64_t b200w_sweep_dist_tiles(int64_t n, int world, int rank, int64_t per, int32_t* out_xy, int64_t capacity);
/* Host only: with more than two ranks the half of the similarity that a rank computes for a peer is ONE rectangle, written
 * to a local staging area by the kernel and pushed into the peer's block by the copy engines.  out[4 * o ..] = {row0, row1,
 * col0, col1} (global gene indices, half open) of the rectangle `rank` owes peer o, zeros where there is none.  Returns 1
 * (staged), 0 (two ranks: the kernel stores into the peer directly) or -1 (bad arguments). */
int b200w_sweep_dist_regions(int64_t n, int world, int rank, int64_t per, int64_t* out /* 4 * world */);
/* statistics of k[P, ldk >= n] as described at b200w_soft_threshold_stats; d_stats P x (4 + 2 nb),
 * d_buckets P x b200w_sft_exp_buckets() x 2 */
int b200w_dev_sft_stats(const double* d_k, int P, int64_t n, int64_t ldk, int nbreaks, double* d_stats,
                        int64_t* d_buckets, int device, void* stream);
int b200w_dev_adjacency(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, int net_type,
                        double power, int64_t row0, int64_t nrows, double* dA, int device,
                        void* stream);
/* The statistics above plus, still on the device, what WGCNA.pickSoftThreshold decides from them
 * (wgcna.py:924-953): d_fit[p] = (R^2 of the first ten-point regression of scaleFreeFitIndex :1039, exactly
 * rounded statistics.mean :930), and d_power_out[0] = the selected power = the smallest of d_powers[P]
 * (ascending, device memory) with R^2 > rsq_cut and mean(k) <= mean_cut, else the power of the largest R^2
 * (:945-953); d_power_out[1] = its index.  Nothing is copied to the host: the *_dp entry points below read
 * the power from device memory, so a chained step never waits for the host. */
int b200w_dev_sft_select(const double* d_k, int P, int64_t n, int64_t ldk, int nbreaks, const double* d_powers,
                         double rsq_cut, double mean_cut, double* d_stats, int64_t* d_buckets,
                         double* d_fit /* P x 2 */, double* d_power_out /* 2 */, int device, void* stream);
int b200w_dev_adjacency_dp(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, int net_type,
                           const double* d_power, int64_t row0, int64_t nrows, double* dA, int device,
                           void* stream);
int b200w_dev_adjacency_from_similarity_dp(double* d_sim_inout, int64_t nrows, int64_t n, const double* d_power,
                                           int device, void* stream);
int b200w_dev_check_adjacency(const double* dA, int64_t n, double* d_stats6, int device,
                              void* stream);

/* TOM on device-resident data, three steps so that a multi-GPU caller can exchange the
 * operand between them:
 *  (1) b200w_dev_tom_prepare: for the row block [row0,row0+nrows) of A: NaN -> 0, diag -> 0,
 *      k = row sums, and the operand the contraction needs --
 *        precision F64: Ac  nrows x ldn float64 (ldn = b200w_padded_n(n)), zero padded
 *        precision I8 : Q   B200W_I8_SLICES planes of nrows x ldn int8 digits + scale[nrows]
 *      written at row offset row0 of the FULL operand buffers (n rows), so that after an
 *      all-gather (or with nrows == n) the buffers hold the whole operand.
 *  (2) [multi-GPU: all-gather operand row blocks and k between ranks]
 *  (3) b200w_dev_tom_compute: out[row block] from the full operand + k. */
#define B200W_I8_SLICES 5
int64_t b200w_padded_n(int64_t n);
/* bytes of the full operand buffer for n genes at the given precision */
int64_t b200w_tom_operand_bytes(int64_t n, int precision);
/* op_rows = rows allocated per operand plane, >= round_up(n, 128) and a multiple of 128 (a
 * multi-GPU caller pads it to world * rows_per_rank so that equal blocks can be all-gathered);
 * rows >= n must be zero (allocate the buffer zeroed).  d_scale and d_k hold op_rows doubles. */
int b200w_dev_tom_prepare(const double* dA_rows /* nrows x n */, int64_t n, int64_t row0,
                          int64_t nrows, int precision, void* d_operand, int64_t op_rows,
                          double* d_scale, double* d_k, int device, void* stream);
int b200w_dev_tom_compute(const double* dA_rows /* nrows x n, original */, const void* d_operand,
                          int64_t op_rows, const double* d_scale, const double* d_k, int64_t n,
                          int64_t row0,
                          int64_t nrows, int tom_type, int tom_denom, int out_mode, int precision,
                          double* d_out /* nrows x n */, int device, void* stream);

/* Step (3) on one GPU with the result delivered to HOST memory while the kernel is still running: finished
 * row blocks are reported through host-mapped flags and copied out on a second stream (int8 precision, whole
 * matrix, out_mode TOM / 1 - TOM / rounded).  d_out is the nrows x n device buffer the kernel writes; returns when
 * out_host is complete. */
int b200w_dev_tom_compute_to_host(const double* dA /* n x n, original */, const void* d_operand, int64_t op_rows,
                                  const double* d_scale, const double* d_k, int64_t n, int tom_type,
                                  int tom_denom, int out_mode, int precision, double* d_out, double* out_host,
                                  int device, void* stream);

/* Step (3) for several GPUs, symmetric and fused with its exchange.  TOM is symmetric, so only the
 * tiles on and above the diagonal are computed, dealt out over the `world` ranks; rank r owns the
 * output rows [r * per, (r + 1) * per) (per % 256 == 0) in ITS buffer peer_out[r] (per x n doubles).
 * Every tile's direct half is stored into the local block, its mirrored half straight into the block of
 * the rank that owns those rows: peer_out[q] for q != rank are PEER-MAPPED pointers (b200w_ipc_open),
 * so the stores cross NVLink from inside the tcgen05 kernel -- no separate collective.  A rank's block
 * is complete once every rank's call has finished (stream-order a collective, or synchronise, after it).
 * int8 precision only.  dA_rows = this rank's rows of the original adjacency. */
int b200w_dev_tom_compute_dist(const double* dA_rows, const void* d_operand, int64_t op_rows,
                               const double* d_scale, const double* d_k, int64_t n, int world, int rank,
                               int64_t per, int tom_type, int tom_denom, int out_mode,
                               void* const* peer_out /* host array of `world` device pointers */,
                               int device, void* stream);

/* The same contraction with its operand exchange overlapped (multi-GPU).  Instead of all-gathering the digit
 * planes before the kernel starts, every rank keeps planes, scale and k in ONE shareable allocation
 * (b200w_dev_malloc of b200w_tom_shared_bytes(op_rows, n): planes at 0, scale at
 * b200w_tom_shared_scale_offset, k right behind it; d_operand / d_scale / d_k of the calls above point into
 * it) and, once all ranks have prepared their rows, calls b200w_dev_tom_pull_planes on a COPY stream: peer by
 * peer (rank + 1, rank + 2, ...) the rows of scale, k and the five planes are copied out of that peer's HBM by
 * the copy engines (peer_base[q] = rank q's allocation, peer-mapped), and after each peer d_plane_flags[q] is set
 * to `epoch` by a stream memory operation.  b200w_dev_tom_compute_dist_gated, launched at once on the compute
 * stream, visits its tiles in the same order -- first those whose column operand is local, then peer rank + 1,
 * ... -- and its TMA producer waits at the first tile of each peer until that flag has reached `epoch` (a
 * wrapping 32-bit counter the caller increments per exchange; d_plane_flags holds 8 words). */
int64_t b200w_tom_shared_scale_offset(int64_t op_rows, int64_t n);
int64_t b200w_tom_shared_bytes(int64_t op_rows, int64_t n);
int b200w_dev_tom_pull_planes(void* d_local, void* const* peer_base, int world, int rank, int64_t per,
                              int64_t op_rows, int64_t n, uint32_t* d_plane_flags /* 8 words, one per rank */,
                              uint32_t epoch, int device, void* copy_stream);
int b200w_dev_tom_compute_dist_gated(const double* dA_rows, const void* d_operand, int64_t op_rows,
                                     const double* d_scale, const double* d_k, int64_t n, int world, int rank,
                                     int64_t per, int tom_type, int tom_denom, int out_mode,
                                     void* const* peer_out, const uint32_t* d_plane_flags, uint32_t plane_epoch,
                                     int device, void* stream);

/* Device buffers that can be shared between the per-GPU processes of one node (plain cudaMalloc, hence
 * exportable), and the CUDA IPC plumbing for them.  handle = 64 bytes (cudaIpcMemHandle_t). */
int b200w_dev_malloc(int64_t bytes, int device, void** out);
int b200w_dev_free(void* p, int device);
int b200w_ipc_export(void* p, unsigned char* handle64);
int b200w_ipc_open(const unsigned char* handle64, int device, void** out);
int b200w_ipc_close(void* p, int device);

/* Host <-> device copies for callers that keep the chain on the device (DeviceNetwork) but hand results to
 * numpy: DMA into / out of a caller buffer (page-locked for the duration of the call unless it already is,
 * e.g. a b200w_host_alloc buffer); the call returns when the copy is complete. */
int b200w_dev_download(void* host_dst, const void* dev_src, int64_t bytes, int device, void* stream);
int b200w_dev_upload(void* dev_dst, const void* host_src, int64_t bytes, int device, void* stream);

/* ----------------------------------------------------------------------- introspection */
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t b200w_launch_count(void);
/* device time in ms of the most recent b200w_dev_tom_compute contraction kernel as measured by
 * CUDA events on its stream when B200W_TIME_KERNELS=1 is set (else -1) */
double b200w_last_kernel_ms(void);

#ifdef __cplusplus
}
#endif
#endif /* B200WGCNA_H */
