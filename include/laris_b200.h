/*
 * laris_b200.h - C ABI of the B200-native LARIS per-cell LR scoring path.
 *
 * The reference (genecell/LARIS) has no FFI/plugin interface: its boundary is
 * the Python call surface laris.tl.prepareLRInteraction / laris.tl.runLARIS
 * (reference laris/tools/__init__.py:29-34, :121-153) and it dispatches the
 * arithmetic to compiled routines inside scikit-learn / scipy / numpy.  Each
 * entry point below replaces one of those dispatches; the comment on each
 * names the reference call site it stands in for.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless it says "host";
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     nothing synchronises the host unless the comment says so;
 *   - return value: 0 = ok, <0 = error (LARIS_E_*); text via laris_last_error();
 *   - scratch memory is taken from the stream-ordered allocator
 *     (cudaMallocAsync/cudaFreeAsync on `stream`), never from the caller's buffers;
 *   - row-major everywhere; "ld" is a row stride in ELEMENTS.
 */
#ifndef LARIS_B200_H
#define LARIS_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define LARIS_API __attribute__((visibility("default")))
#else
#define LARIS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LARIS_OK 0
#define LARIS_E_ARG (-1)     /* bad argument */
#define LARIS_E_CUDA (-2)    /* CUDA runtime error */
#define LARIS_E_LIMIT (-3)   /* size exceeds a compiled limit */

#define LARIS_ABI_VERSION 1

LARIS_API int laris_abi_version(void);
/* Number of kernels this library has launched in the process so far (bench.py's gpu_launches). */
LARIS_API int64_t laris_launch_count(void);
/* Thread-local text of the last error returned on this thread (host pointer). */
LARIS_API const char* laris_last_error(void);

/* ---------------------------------------------------------------- K1 ----
 * Exact Euclidean kNN graph on a uniform grid, rows ascending by distance.
 * Replaces sklearn.neighbors.kneighbors_graph(mode='distance') at
 * laris/tools/__init__.py:83-88 (include_self=1) and laris/tools/_utils.py:385-390,
 * :847-852 (include_self=0: query k+1, drop the row's own index, or the first
 * entry when the row id is absent).  d^2 = ((0+dx*dx)+dy*dy[+dz*dz]) in fp64
 * without FMA, then IEEE sqrt; ties ordered by (d^2, index).
 *   xy        [n*dim] fp64, dim in {2,3}
 *   out_idx   [n*k]   int32 neighbour ids (the CSR 'indices'; indptr = i*k)
 *   out_dist  [n*k]   fp64 distances (the CSR 'data' before the weight transform)
 *   out_order [n]     int32 spatial processing order (grid-cell sorted), may be NULL
 * Requires k(+1 if !include_self) <= 64 and n >= k+1.  Fully asynchronous: the grid shape is derived on
 * the device from the bounding box (no host round trip).  Non-finite coordinates cannot be reported
 * through the return code without a synchronisation: every out_idx is then -1 and every out_dist NaN.
 */
LARIS_API int laris_knn_graph(const double* xy, int64_t n, int dim, int k, int include_self,
                    int32_t* out_idx, double* out_dist, int32_t* out_order, void* stream);

/* ---------------------------------------------------------------- K1r ---
 * Radius graph (north-star item 1; EXTENSION - the reference only ever calls kneighbors_graph,
 * laris/tools/__init__.py:83, laris/tools/_utils.py:385,847).  Oracle:
 * sklearn.neighbors.radius_neighbors_graph(mode='distance', include_self=...): every j with
 * d^2(i,j) <= radius^2 (d^2 accumulated as in laris_knn_graph), self kept only when include_self.
 * Two passes over the same uniform grid (cell edge >= radius):
 *   count: counts [n] int64 row degrees (+ optional spatial processing order [n] int32);
 *          the caller scans them into indptr (laris_exclusive_scan_i64)
 *   fill:  indices [nnz] int32 column-sorted within a row (canonical CSR), dist [nnz] fp64
 * Both passes are asynchronous (grid shape derived on the device); non-finite coordinates give empty rows. */
LARIS_API int laris_radius_graph_count(const double* xy, int64_t n, int dim, double radius, int include_self,
                                       int64_t* counts, int32_t* out_order, void* stream);
LARIS_API int laris_radius_graph_fill(const double* xy, int64_t n, int dim, double radius, int include_self,
                                      const int64_t* indptr, int32_t* indices, double* dist, void* stream);
/* Variable-degree CSR rows -> fixed-degree rows idx/w_out [n*kmax] (kmax >= max degree): padding edges point
 * at the row itself with weight 0, so every weighted neighbour sum downstream is unchanged. */
LARIS_API int laris_csr_pad_rows(const int64_t* indptr, const int32_t* indices, const float* w, int64_t n, int kmax,
                                 int32_t* idx, float* w_out, void* stream);

/* w = 1/exp(d/s), s = sigma if sigma > 0 else mean(all m distances)/2.
 * Replaces the .data transforms at laris/tools/__init__.py:89 and
 * laris/tools/_utils.py:392, :853.  Either output may be NULL.
 * scale_out (device, 1 double, may be NULL) receives s. */
LARIS_API int laris_decay_weights(const double* dist, int64_t m, double sigma, double* w64, float* w32,
                        double* scale_out, void* stream);

/* ------------------------------------------------------- gene selection --
 * Restrict a CSR expression matrix to the genes the LR table uses and pack it
 * for the diffusion kernel (part of the fancy-index gathers at
 * laris/tools/__init__.py:100, :113-114).
 *   indptr [n+1] int64, indices [nnz] int32, data [nnz] fp32
 *   gene_map [G] int32: new column id in [0,G_u) or -1 to drop
 * Phase 1 writes per-row kept counts to out_counts[n] (int64); the caller
 * exclusive-scans them (laris_exclusive_scan_i64) into out_indptr[n+1];
 * phase 2 writes packed entries {int32 col, fp32 val} as int64 words,
 * dropping explicitly stored zeros.
 */
LARIS_API int laris_csr_select_count(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                           const int32_t* gene_map, int64_t* out_counts, void* stream);
LARIS_API int laris_csr_select_fill(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                          const int32_t* gene_map, const int64_t* out_indptr, int64_t* out_packed,
                          void* stream);
/* Single-pass variant of the two calls above: the kept entries of row r are written to
 * out_packed[indptr[r] ...) - every row keeps its offset in the INPUT matrix, so neither a count pass nor a
 * scan nor a host read of the total is needed - and row_end[r] (int64 [n]) marks where they stop.
 * out_packed needs nnz(X) slots.  Consumers take (row_start = indptr, row_end): laris_diffuse_lr_score_rows,
 * laris_onehot_contract_rows. */
LARIS_API int laris_csr_select_rows(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                          const int32_t* gene_map, int64_t* out_packed, int64_t* row_end, void* stream);
/* out[0]=0, out[i+1]=sum(in[0..i]); in [n], out [n+1]; may alias in == out+1 is NOT allowed. */
LARIS_API int laris_exclusive_scan_i64(const int64_t* in, int64_t n, int64_t* out, void* stream);

/* ------------------------------------------------------------- K2+K3 ----
 * Fused neighbour-aggregated expression + per-cell LR score:
 *   D[i,g] = sum_j w[i,j] * Xu[idx[i,j], g]
 *   S[i,p] = D[i,lig[p]] * D[i,rec[p]] * [Xu[i,lig[p]] != 0 || Xu[i,rec[p]] != 0]
 * Replaces the SpGEMM at laris/tools/__init__.py:91-92, the row gathers and
 * product at :100 and the expression mask at :110-117.  D is never written
 * to HBM.
 *   xu_indptr [n+1] int64, xu_packed [nnz] from laris_csr_select_fill
 *   idx [n*k] int32, w [n*k] fp32 (K1 outputs; include_self graph)
 *   order [n] int32 processing order or NULL
 *   lig, rec [P] int32 in [0,G_u)
 *   S [n*ldS] fp32 (ldS >= P, ldS % 4 == 0, base 16-byte aligned); columns [P,ldS) are zeroed
 *   row_nnz [n] int64 number of non-zero scores per cell (scan it for the CSR indptr), may be NULL
 */
LARIS_API int laris_diffuse_lr_score(const int64_t* xu_indptr, const int64_t* xu_packed, int64_t n, int32_t G_u,
                           const int32_t* idx, const float* w, int k, const int32_t* order,
                           const int32_t* lig, const int32_t* rec, int32_t P,
                           float* S, int64_t ldS, int64_t* row_nnz, void* stream);

/* Dense-panel variant of the same fused path, for expression panels dense enough (>= ~8 % non-zeros
 * over the LR genes) that dense fp32 rows are cheaper to gather than packed sparse rows.  Same
 * reference lines as above (laris/tools/__init__.py:91-92, :100, :110-117).
 *
 * laris_csr_select_dense: restrict the CSR matrix to the LR genes and write them as dense rows
 *   Xd [n*ldX] fp32, ldX % 128 == 0, ldX > (largest column id in gene_map); slots without a gene
 *   are zero.  neg_flag [1] int32 (may be NULL) is set to 1 if any kept value is negative.
 * laris_diffuse_lr_score_dense: one warp owns min(3, 32/k) consecutive cells of `order`, loads each
 *   distinct neighbour row once (128-bit loads) and accumulates D for all of them in registers.
 *   neg_flag from the call above (NULL = unknown: the slower bitmap path is taken); the other
 *   arguments as in laris_diffuse_lr_score.  Result: identical structure, values equal up to the
 *   fp32 summation order over neighbours. */
LARIS_API int laris_csr_select_dense(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                           const int32_t* gene_map, float* Xd, int64_t ldX, int32_t* neg_flag, void* stream);
LARIS_API int laris_diffuse_lr_score_dense(const float* Xd, int64_t ldX, const int32_t* neg_flag, int64_t n,
                                 int32_t G_u, const int32_t* idx, const float* w, int k, const int32_t* order,
                                 const int32_t* lig, const int32_t* rec, int32_t P,
                                 float* S, int64_t ldS, int64_t* row_nnz, void* stream);

/* Multi-subunit complexes (north-star item 3; EXTENSION - the reference has no complex handling and scores
 * one LR row per subunit pair, docs/notebooks/04_LARIS_tutorial.html:743,767).  Same kernels and arguments as
 * laris_diffuse_lr_score / laris_diffuse_lr_score_dense, plus a complex table: complex c owns the packed
 * column cx_vcol[c] (a column no gene maps to) and is aggregated from the subunit columns
 * cx_cols[cx_ptr[c] .. cx_ptr[c+1]) after the diffusion: rule 0 = min over subunits, rule 1 = geometric mean
 * of max(v,0); it is "expressed" in a cell only if every subunit is.  lig/rec may name virtual columns. */
LARIS_API int laris_diffuse_lr_score_complex(const int64_t* xu_indptr, const int64_t* xu_packed, int64_t n, int32_t G_u,
                           const int32_t* idx, const float* w, int k, const int32_t* order,
                           const int32_t* lig, const int32_t* rec, int32_t P,
                           float* S, int64_t ldS, int64_t* row_nnz, const int32_t* cx_ptr, const int32_t* cx_cols,
                           const int32_t* cx_vcol, int32_t n_complex, int32_t rule, void* stream);
/* laris_diffuse_lr_score_complex over rows given as [row_start[i], row_end[i]) ranges of xu_packed (the layout
 * laris_csr_select_rows produces; for a compact CSR pass indptr and indptr + 1). */
LARIS_API int laris_diffuse_lr_score_rows(const int64_t* row_start, const int64_t* row_end, const int64_t* xu_packed,
                           int64_t n, int32_t G_u, const int32_t* idx, const float* w, int k, const int32_t* order,
                           const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS, int64_t* row_nnz,
                           const int32_t* cx_ptr, const int32_t* cx_cols, const int32_t* cx_vcol, int32_t n_complex,
                           int32_t rule, void* stream);
LARIS_API int laris_diffuse_lr_score_dense_complex(const float* Xd, int64_t ldX, const int32_t* neg_flag, int64_t n,
                           int32_t G_u, const int32_t* idx, const float* w, int k, const int32_t* order,
                           const int32_t* lig, const int32_t* rec, int32_t P, float* S, int64_t ldS, int64_t* row_nnz,
                           const int32_t* cx_ptr, const int32_t* cx_cols, const int32_t* cx_vcol, int32_t n_complex,
                           int32_t rule, void* stream);

/* HOST helper (all pointers are host memory, no device work): number the Xu columns so that
 * laris_diffuse_lr_score's 32-lane gathers of D[lig[p]] / D[rec[p]] are free of shared-memory
 * bank conflicts.  lig, rec [P] are gene indices in [0,G_u); column_of [G_u] receives the column
 * of each gene (injective), *n_columns the column range (a multiple of 32, >= G_u).  Feed
 * column_of through gene_map of laris_csr_select_* and use column_of[lig], column_of[rec] and
 * G_u = *n_columns in laris_diffuse_lr_score.  Any injective numbering is CORRECT; this one is
 * the fast one. */
LARIS_API int laris_lr_column_layout(const int32_t* lig, const int32_t* rec, int32_t P, int32_t G_u,
                           int32_t* column_of, int32_t* n_columns);

/* Dense S -> CSR (the reference's lr_adata.X, laris/tools/__init__.py:104,117):
 * indptr [n+1] int64 = exclusive scan of row_nnz; writes column-sorted
 * indices [nnz] int32 and data [nnz] fp32 holding exactly the non-zero scores. */
LARIS_API int laris_csr_compact(const float* S, int64_t ldS, int64_t n, int32_t P, const int64_t* indptr,
                      int32_t* out_indices, float* out_data, void* stream);
/* CSR -> dense S (for a caller-supplied lr_adata.X); zero-fills then scatters. */
LARIS_API int laris_csr_to_dense(const int64_t* indptr, const int32_t* indices, const float* data, int64_t n,
                       int32_t P, float* S, int64_t ldS, void* stream);

/* ---------------------------------------------------------------- K4 ----
 * Per-pair spatial statistics in ONE pass over S, T never materialised:
 *   T[i,p] = sum_j w[i,j] * S[idx[i,j], p]   (/ sum_j w[i,j] if l1_normalise)
 *   out[0,p]=sum_i S*T  out[1,p]=sum_i S^2  out[2,p]=sum_i T^2
 *   out[3,p]=sum_i S    out[4,p]=#{i: S != 0}
 * Replaces `genexcell @ cellxcell.T` + _rowwise_cosine_similarity at
 * laris/tools/__init__.py:468-470 / laris/tools/_utils.py:81-101 (observed graph,
 * l1_normalise=0) and normalize + SpMM + cosine at laris/tools/_utils.py:512-516
 * (random graph, l1_normalise=1; duplicate (row,col) edges add, as in the
 * reference's COO->CSR).  Sums 3-4 also feed n_cells_by_counts
 * (laris/tools/__init__.py:493-504) and the (mean, variance) features of
 * laris/tools/_utils.py:1081-1088.  Accumulation fp64, deterministic.
 *   out [5*P] fp64
 */
LARIS_API int laris_spatial_stats(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* idx,
                        const float* w, int k, int l1_normalise, const int32_t* order, double* out,
                        void* stream);

/* Random-neighbour graph of one null repeat, device RNG (Philox4x32-10 +
 * Feistel permutation; spec in laris_b200/csrc/rng.cuh, CPU twin oracle/rng.py):
 *   cols[e] uniform in [0,n),  wout[e] = w[perm(e)]  for e in [0, n*k)
 * Replaces np.random.choice + np.random.shuffle at laris/tools/_utils.py:436-442. */
LARIS_API int laris_null_graph_philox(int64_t n, int k, const float* w, int32_t repeat, uint64_t seed,
                            int32_t* cols, float* wout, void* stream);
/* wout[e] = w[perm[e]] for a host-replayed numpy permutation (rng="numpy"). */
LARIS_API int laris_gather_f32(const float* w, const int64_t* perm, int64_t m, float* wout, void* stream);

/* ---------------------------------------------------------------- K5 ----
 * Neighbourhood composition N[i,a] = sum of w[i,j] over neighbours j of type a
 * (cluster_mat @ cellxcell.T, laris/tools/_utils.py:839-856), cell-major.
 *   labels [n] int32 in [0,C); Ncomp [n*C] fp32 */
LARIS_API int laris_neighborhood_composition(const int32_t* labels, const int32_t* idx, const float* w, int64_t n,
                                   int k, int32_t C, float* Ncomp, void* stream);

/* Cell-type-pair contraction (the dense contraction over cells):
 *   num[p, a*C+b] = sum_i S[i,p] * N[i,a] * N[i,b]
 *   qnorm2[a*C+b] = sum_i (N[i,a]*N[i,b])^2
 * Replaces _pairwise_row_multiply + sklearn cosine_similarity at
 * laris/tools/_utils.py:859-866.  impl: 0 = tcgen05 tensor-core kernel
 * (fp32 inputs split three ways into bf16 hi/mid/lo, six MMAs per 16-cell K step, fp32 TMEM accumulation
 * in 256-cell chunks, fixed-order fp64 combine; needs C <= 64), 1 = fp32 CUDA-core kernel (any C).
 *   order [n] int32 or NULL: cell processing order for impl=0 (a spatially coherent order lets the
 *   operand builder skip the all-zero cell-type blocks); the result does not depend on it beyond
 *   fp32 summation order.
 *   num [P*C*C] fp64, qnorm2 [C*C] fp64 */
LARIS_API int laris_celltype_pair_contract(const float* S, int64_t ldS, int64_t n, int32_t P, const float* Ncomp,
                                 int32_t C, const int32_t* order, double* num, double* qnorm2, int impl, void* stream);

/* cnt[c*P + p] = #{i : labels[i]==c and S[i,p] != 0}.  Replaces
 * _compute_avg_expression on the binarised scores (laris/tools/_utils.py:1325-1337,
 * :334-338); frac = cnt / n_c on the host.  counts [C*P] uint64; any C (cells are grouped by type on the
 * device and counted in registers, integer atomics only where the type changes); labels outside [0,C) are skipped. */
LARIS_API int laris_celltype_nonzero_count(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* labels,
                                 int32_t C, uint64_t* counts, void* stream);
/* The same counts (same reference lines) from the packed rows of S that laris_scores_pack wrote (sp_ptr [n+1], entries
 * {int32 column, fp32 bits}, zero-valued padding; entries beyond `capacity` are not read): 8 bytes per non-zero score
 * instead of 4P bytes per cell.  P * 4 bytes of shared memory per CTA. */
LARIS_API int laris_celltype_nonzero_count_packed(const int64_t* sp_ptr, const void* entries, int64_t capacity, int64_t n,
                                 int32_t P, const int32_t* labels, int32_t C, uint64_t* counts, void* stream);
/* One-hot contraction straight from a packed CSR (laris_csr_select_fill layout):
 * sum[c*G_u+g] = sum of Xu over cells of type c, cnt = non-zero count, colsq[g] =
 * sum_i Xu[i,g]^2.  These are the numerators / norms of the cosine between a
 * gene and the one-hot group indicator in cosg.cosg (call site
 * laris/tools/_utils.py:580-588) and its low-expression mask.  fp64 outputs. */
LARIS_API int laris_onehot_contract_csr(const int64_t* indptr, const int64_t* packed, int64_t n, int32_t G_u,
                              const int32_t* labels, int32_t C, double* sum, double* cnt, double* colsq,
                              void* stream);
/* The same over [row_start[i], row_end[i]) ranges (see laris_csr_select_rows). */
LARIS_API int laris_onehot_contract_rows(const int64_t* row_start, const int64_t* row_end, const int64_t* packed, int64_t n,
                              int32_t G_u, const int32_t* labels, int32_t C, double* sum, double* cnt, double* colsq,
                              void* stream);

/* ------------------------------------------ K5, key-gather ("sparse B") ----
 * The same numerators as laris_celltype_pair_contract for neighbourhoods that hold few cell types (k <= ~10
 * neighbours: 1-3 non-zero N[i,a] per cell, so the n x C^2 operand is > 99 % zeros).  Q is kept sparse with the key
 * (a <= b) as column; each key's cells are gathered (contiguous row pieces, 128-bit streaming loads) and q*S is
 * accumulated in fp64 - no operand split, no TMEM truncation.  HBM bytes: 4*P per entry, i.e. S is read once per key
 * a cell takes part in.  Two calls, because the entry total sizes the scratch and picks the implementation:
 *   laris_celltype_pair_keys:  key_count [C*C+1] int32 <- number of cells per key a*C+b (a <= b), rest 0;
 *                              the caller sums it (one tiny device->host read) to get total_entries
 *   laris_celltype_pair_contract_sparse: num [P*C*C] fp64 (both (a,b) and (b,a) written); key_count is consumed
 * laris_pair_profile_norm: qnorm2 [C*C] = sum_i (N[i,a] N[i,b])^2 (what laris_celltype_pair_contract also returns).
 * C <= 64, n < 2^31.  Entry order inside a key follows device atomics: fp64 sums can differ in the last bits
 * between runs. */
LARIS_API int laris_celltype_pair_keys(const float* Ncomp, int64_t n, int32_t C, int32_t* key_count, void* stream);
LARIS_API int laris_celltype_pair_contract_sparse(const float* S, int64_t ldS, int64_t n, int32_t P, const float* Ncomp,
                                                  int32_t C, int32_t* key_count, int64_t total_entries, double* num,
                                                  void* stream);
LARIS_API int laris_pair_profile_norm(const float* Ncomp, int64_t n, int32_t C, double* qnorm2, void* stream);

/* ------------------------------------ label-permutation null (extension) ----
 * North-star item 4 names a "label permutation null"; the reference has none (its nulls are the random graph,
 * laris/tools/_utils.py:436-446, and the background-pair resampling, :1544-1570).  Extension with its own CPU oracle
 * (oracle/restate.py: label_permutation_null): permutation t relabels the cells with a keyed bijection of [0, n)
 * (the Feistel network of the K4b weight shuffle, key stream 0x40000000 + t, bit-exact numpy twin in oracle/rng.py),
 * the neighbourhood composition and the P x C^2 co-localisation cosines are recomputed, and
 * ge[p,(a,b)] counts cos_t >= cos_observed.  Relabelled neighbourhoods are maximally mixed, i.e. the B operand of the
 * contraction is dense: this is the workload the tcgen05 kernel (laris_celltype_pair_contract impl 0) is for.
 *   laris_permute_labels: out[i] = labels[pi_t(i)]
 *   laris_count_exceed:   ge[e] += val[e] >= obs[e]   (fp64 tables of `total` entries, int32 counts) */
LARIS_API int laris_permute_labels(const int32_t* labels, int64_t n, int32_t permutation, uint64_t seed, int32_t* out,
                                   void* stream);
LARIS_API int laris_count_exceed(const double* obs, const double* val, int64_t total, int32_t* ge, void* stream);

/* ---------------------------------------------------------------- K6 ----
 * Exceedance counts of the background-resampling null
 * (laris/tools/_utils.py:1518-1596).  Row r = (group g, pair s), s over ALL pairs:
 *   obs = table[g*ldt + s]; null_t = table[g*ldt + bg[s*nb + draw_t]]
 *   ge[r] += #{t: null_t >= obs}, nz[r] += #{t: null_t > 0}, ge_nz[r] += #{both}
 * (pairs outside the ranked list carry 0.0 in the table, reference :1563, and
 * are skipped as rows through active[s] == 0).
 * rng_mode 0: Philox draws keyed by (seed, global row id, t) with global row id
 *   = (g0+g)*P_global + (slot_gid ? slot_gid[s] : s) and t in [t0, t0+n_perm) -
 *   independent of how pairs / groups / permutations are sharded;
 * rng_mode 1: draws [n_groups*n_slots*n_perm] uint8 supplied by the caller
 *   (host replay of the reference's numpy stream).
 *   table [n_groups*ldt] fp64, bg [n_slots*nb] int32, active [n_slots] uint8 or NULL
 *   ge, nz, ge_nz [n_groups*n_slots] int32, accumulated into (zero them first); nz and ge_nz may both be
 *   NULL when only the standard p-value is wanted (they feed the conditional variant, :1573-1590)
 * Philox draws: for nb <= 64 each 32-bit word w gives three draws, the first three base-nb digits of w/2^32
 * (idx = (w*nb) >> 32; w = (w*nb) mod 2^32), so permutation t uses word (t/3)%4 of call t/12; else one per word.
 */
LARIS_API int laris_pvalue_counts(const double* table, int64_t ldt, int32_t n_groups, int32_t n_slots,
                        const int32_t* bg, int32_t nb, const uint8_t* active, const int64_t* slot_gid,
                        int64_t P_global, int32_t g0, int32_t t0, int32_t n_perm, int rng_mode,
                        uint64_t seed, const uint8_t* draws, int32_t* ge, int32_t* nz, int32_t* ge_nz,
                        void* stream);

/* ---------------------------------------------------------------- K7 ----
 * Benjamini-Hochberg per group (statsmodels multipletests 'fdr_bh' at
 * laris/tools/_utils.py:1636-1646): p [n_groups*len] fp64, sel [n_groups*len]
 * uint8 marks the tested rows; out fdr [n_groups*len] fp64 (1.0 where !sel). */
LARIS_API int laris_bh_fdr(const double* p, const uint8_t* sel, int32_t n_groups, int32_t len, double* fdr,
                 void* stream);

/* ------------------------------------------- step-2 tables (T1..T5) ----
 * The P x C^2-sized fp64 tables between the S-sized passes and K6/K7 stay on the device; these entry points
 * replace the reference's pandas / numpy / sklearn arithmetic on them.  L = C*C, group g = sender*C + receiver.
 *
 * T1 (laris/tools/_utils.py:862-880): cos[p,l] = num[p,l] / (sqrt(ss[p]) * sqrt(qn2[l])) (0 where the
 * denominator is 0) - sklearn cosine_similarity of the score columns against the cell-type-pair profiles - and
 * reg = cos^2 / ((1-mu) cos^2 + mu * sum_l cos^2) * cos (mu == 1: / sum_l cos^2), exact zeros staying zero.
 * num, cos_out (may be NULL), reg_out: [P*L] fp64; ss [P], qn2 [L] fp64. */
LARIS_API int laris_pair_cosine_regularise(const double* num, const double* ss, const double* qn2, int32_t P, int32_t L,
                                           double mu, double* cos_out, double* reg_out, void* stream);
/* T2: stand-in of cosg.iqrLogNormalize (call sites laris/tools/_utils.py:596, :975; the package is not vendored
 * by the reference - PARITY UNPINNED).  Per column of table [rows*cols]: scale = Q(q_upper) - Q(q_lower) of the
 * positive entries (numpy 'linear' quantiles; their maximum when that difference is not positive),
 * out = log1p(max(x,0)/scale) / max over the column, i.e. values in [0,1]; a column without positive entries is 0. */
LARIS_API int laris_column_iqr_log(const double* table, int64_t rows, int64_t cols, double q_upper, double q_lower,
                                   double* out, void* stream);
/* T3a (laris/tools/_utils.py:1351-1406): inter[t,s,r] = (frac[p,s]*frac[p,r]) * ((spec_l[t,s]*spec_r[t,r]) * factor[t]),
 * p = pair_id[t], frac[p,c] = cnt[c*P+p] / n_c[c] (K5b counts).  spec_l, spec_r [Pt*C], factor [Pt], n_c [C] fp64;
 * cnt [C*P] int64; pair_id [Pt] int32; inter [Pt*C*C] fp64. */
LARIS_API int laris_interaction_scores(const double* spec_l, const double* spec_r, const double* factor,
                                       const int64_t* cnt, const double* n_c, const int32_t* pair_id, int32_t Pt,
                                       int32_t P, int32_t C, double* inter, void* stream);
/* T3b (:1426-1437): out2[0] = mean of the min(k, m) largest of values[0..m), out2[1] = 0.1 / out2[0] (1 when the
 * mean is not positive).  Exact (radix select); k <= 4096. */
LARIS_API int laris_topk_mean(const double* values, int64_t m, int32_t k, double* out2, void* stream);
/* T3c (:1463-1489): flat[t,g] = (inter[t,g] * rescale2[1]) * ctc[pair_id[t]*L + g], values below score_threshold
 * -> 0; also scattered into the group-major table [L*P] (zero elsewhere) and active [P] that K6 reads. */
LARIS_API int laris_finalize_scores(const double* inter, const double* rescale2, const double* ctc,
                                    const int32_t* pair_id, int32_t Pt, int32_t P, int32_t L, double score_threshold,
                                    double* flat, double* table, uint8_t* active, void* stream);
/* T5a (:1573-1594, :1617-1634): p = (ge+1)/(n_perm+1), or the conditional (ge_nz+1)/(nz+1) where the score is
 * positive and 1 elsewhere; p_round = numpy.round(p, 6); sel = active pair and (score > prefilter_threshold when
 * prefilter).  All arrays group-major [L*P]. */
LARIS_API int laris_pvalues_from_counts(const double* table, const uint8_t* active, const int32_t* ge, const int32_t* nz,
                                        const int32_t* ge_nz, int32_t L, int32_t P, int32_t n_perm, int conditional,
                                        int prefilter, double prefilter_threshold, double* p_out, double* p_round,
                                        uint8_t* sel, void* stream);
/* T5b (:1648-1658): rows layout [Pt*L] of the group-major p / fdr tables; fdr clipped to <= 1, NaN -> 1,
 * nlog = max(-log10(fdr + 1e-10), 0). */
LARIS_API int laris_gather_result_rows(const double* p_tab, const double* fdr_tab, const int32_t* pair_id, int32_t Pt,
                                       int32_t P, int32_t L, double* p_rows, double* fdr_rows, double* nlog_rows,
                                       void* stream);
/* T6a (laris/tools/_utils.py:1408-1422, :1472): the row order of the reference's long result table.  The reference
 * melts the [pair][receiver][sender] cube (pandas melt: pair-major, receiver-major, sender fastest) and sorts it by
 * the pre-weight score and then by the final score, both descending.  order[j] = element (pair*C + sender)*C +
 * receiver of the [Pt][C][C] tables `inter` / `flat` at row j: the rows with a positive final score stably sorted by
 * both keys, every other row behind them in melt order.  *exotic (device int32) is set when a final score is negative
 * or NaN (only possible with exotic thresholds; the caller then orders on the host).  Two stable 64-bit LSD radix
 * sorts (cub::DeviceRadixSort); Pt*C*C < 2^31. */
LARIS_API int laris_result_row_order(const double* inter, const double* flat, int32_t Pt, int32_t C, int32_t* order,
                                     int32_t* exotic, void* stream);
/* T6b (:1408-1420, :1660-1688): the result table's numeric columns and per-row codes in table order, one pass:
 * score_o[j] = flat[order[j]], likewise p / fdr / nlog / extra (null sources are skipped), pair_o / send_o / recv_o =
 * the row's position in the ranked pair list, sender and receiver category codes. */
LARIS_API int laris_result_rows_gather(const int32_t* order, int64_t total, int32_t C, const double* flat,
                                       const double* p_rows, const double* fdr_rows, const double* nlog_rows,
                                       const double* extra, double* score_o, double* p_o, double* fdr_o, double* nlog_o,
                                       double* extra_o, int32_t* pair_o, int32_t* send_o, int32_t* recv_o, void* stream);

/* ------------------------------------------------- K4a + K4b batched ----
 * Observed statistics and n_repeats random-graph repeats of the same score matrix in one call
 * (laris/tools/__init__.py:468-481, laris/tools/_utils.py:496-518), out [(obs ? 1 : 0) + n_repeats][5][P] fp64 with the
 * rows of laris_spatial_stats (dot, sum S^2, sum T^2, sum S, nnz); row 0 is the observed graph when idx != NULL.
 *   observed graph (idx, w [n*k], order [n] or NULL): for k <= 15 the distinct neighbour rows of every 128-cell tile
 *     of the processing order are staged in shared memory by the TMA engine (cp.async.bulk + mbarrier, double
 *     buffered) and read from there, so DRAM / L2 see each row piece ~1.6 times instead of 1+k times;
 *   random graphs (null_idx int32, null_w fp32, each [n_repeats][n][k]; rows are L1-normalised, duplicates add):
 *     up to three repeats share one pass, so S[i,:] and its sums are read / formed once per batch;
 *     (1/RB + k) * 4*n*P bytes per repeat, the HBM-bound pass of the path. */
LARIS_API int laris_spatial_stats_batched(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* idx,
                                          const float* w, int k, const int32_t* order, const int32_t* null_idx,
                                          const float* null_w, int32_t n_repeats, double* out, void* stream);

/* --------------------------------------- K4a + K4b over packed score rows ----
 * Same statistics as laris_spatial_stats_batched (laris/tools/__init__.py:468-481, laris/tools/_utils.py:496-518; the
 * reference keeps lr_adata.X as a scipy CSR matrix and runs both products as sparse matrix products), for score matrices
 * that are mostly zeros: the neighbour rows are gathered from a PACKED copy of S - 8 bytes per non-zero instead of
 * 4*P bytes per row.
 *   laris_scores_pack: sp_ptr [n+1] int64 = exclusive scan of the per-row non-zero counts (row_nnz of
 *     laris_diffuse_lr_score*) ROUNDED UP to multiples of 32; entries [sp_ptr[n]] int64 = {int32 column, fp32 bits} of
 *     the non-zeros of every row (dealt over the shared-memory banks, not ascending), then padding {ldS + j, 0}.
 *   laris_spatial_stats_sparse: out as laris_spatial_stats_batched; a warp adds the k packed neighbour rows of a cell
 *     into a shared-memory slab in neighbour order (the same fma chain per element as the dense kernels, zeros skipped),
 *     column-owner threads fold the slabs and the cell's own dense row into the sums; up to 4 graphs share one read of S.
 *     `capacity` / `n_entries` = length of `entries`: normally sp_ptr[n]; a caller that sizes the array from an estimate
 *     (to avoid reading sp_ptr[n] back) may pass less - rows that do not fit are skipped by both calls and the caller
 *     must discard the result once it knows sp_ptr[n] > capacity.  ldS <= laris_spatial_stats_sparse_max_ld(). */
LARIS_API int32_t laris_spatial_stats_sparse_max_ld(void);
LARIS_API int laris_scores_pack(const float* S, int64_t ldS, int64_t n, int32_t P, const int64_t* sp_ptr, int64_t* entries,
                                int64_t capacity, void* stream);
LARIS_API int laris_spatial_stats_sparse(const float* S, int64_t ldS, int64_t n, int32_t P, const int64_t* sp_ptr,
                                         const int64_t* entries, int64_t n_entries, const int32_t* idx, const float* w, int k,
                                         const int32_t* order, const int32_t* null_idx, const float* null_w,
                                         int32_t n_repeats, double* out, void* stream);

/* ------------------------------------------------------- per-cell QC ----
 * Per-cell metrics of the score matrix that scanpy.pp.calculate_qc_metrics writes to lr_adata.obs at
 * laris/tools/__init__.py:493-496: total[i] = sum_p S[i,p], nnz[i] = #{p: S[i,p] != 0} and, for each N of
 * top_n (HOST array, n_top <= 4), top_sum[i*n_top + t] = sum of the N largest scores of the cell
 * (= total when N >= P); pct_counts_in_top_N_genes = 100 * top_sum / total. */
LARIS_API int laris_row_qc(const float* S, int64_t ldS, int64_t n, int32_t P, const int32_t* top_n, int32_t n_top,
                           double* total, int64_t* nnz, double* top_sum, void* stream);

/* ------------------------------------------------- laris.pp helpers ----
 * Device versions of the three helpers the reference re-exports as laris.pp
 * (laris/preprocessing/__init__.py:28-38).
 * out[r] = <A_r,B_r>/(|A_r||B_r|), 0 when the denominator is 0
 * (_rowwise_cosine_similarity, laris/tools/_utils.py:81-101); dense fp32 inputs. */
LARIS_API int laris_rowwise_cosine(const float* A, const float* B, int64_t rows, int64_t cols, int64_t lda,
                                   int64_t ldb, double* out, void* stream);
/* out[(a*C+b)*n + i] = M[a*n+i] * M[b*n+i]  (_pairwise_row_multiply,
 * laris/tools/_utils.py:208-245); M [C*n] fp32, out [C*C*n] fp32. */
LARIS_API int laris_pairwise_row_multiply(const float* M, int32_t C, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LARIS_B200_H */
