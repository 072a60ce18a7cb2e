/*
 * qforest_b200.h -- C ABI of the B200-native quantile-forest predict path.
 *
 * Drop-in boundary for scikit-garden's
 *     RandomForestQuantileRegressor / ExtraTreesQuantileRegressor .predict(X, quantile)
 * The reference has no FFI of its own on this path (it is one Python process calling
 * numpy and scikit-learn's Cython); each entry point below names the reference
 * interface it replaces (paths relative to the scikit-garden tree, "sklearn:" =
 * scikit-learn 1.9.0).  The Python host (scikit-garden_b200/) binds these with ctypes;
 * INTEGRATION.md shows the stub a scikit-garden maintainer would add.
 *
 * Conventions
 *   - plain C types, no torch / numpy types; the library never owns caller memory
 *   - every function returns QF_OK (0) or a negative qf_status; the message for the
 *     last failure on the calling thread is returned by qf_last_error()
 *   - nothing throws across the ABI
 *   - "_dev" pointers are device pointers on the forest's device, "_host" are host
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream)
 *   - device entry points are asynchronous w.r.t. the host unless stated otherwise
 *   - a qf_forest handle may be used from several host threads at once (each with its own stream,
 *     workspace and output): per-call state lives in the call.  qf_forest_predict_quantiles_host calls
 *     on one handle share its pipeline buffers and run one after the other; the "last call" counters
 *     report whichever call returned last; qf_forest_set_option is not meant to race with calls
 *   - there is NO CPU fallback: device entry points fail with QF_ERR_CUDA when no
 *     usable sm_100 device is present
 */
#ifndef QFOREST_B200_H
#define QFOREST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QF_ABI_VERSION 2

typedef enum qf_status {
    QF_OK = 0,
    QF_ERR_INVALID = -1,     /* bad argument (NULL, negative size, q outside [0,100], ...) */
    QF_ERR_CUDA = -2,        /* a CUDA runtime call failed; see qf_last_error()            */
    QF_ERR_LIMIT = -3,       /* forest does not fit the packed layout (see qf_pack_tree)   */
    QF_ERR_WORKSPACE = -4    /* caller workspace smaller than qf_forest_workspace_bytes()  */
} qf_status;

/* Arithmetic modes of qf_forest_predict_quantiles. */
#define QF_MODE_EXACT 0   /* bit-identical to the reference's float32 op order            */
#define QF_MODE_FAST  1   /* histogram selection on fixed-point weights; <= 1e-4 relative to
                             the reference.  Forests whose leaf weights are too small for 32-bit
                             fixed point (leaves of hundreds of members) take the exact kernels.
                             Option "select_impl" = 3 runs it on compact leaf lists (32-byte octets of
                             rank | bootstrap count words, built on the device on first use) with
                             asynchronous copies, "select_impl" = 4 with one warp per row and 256-bit
                             loads (no staging); the default (0) reads the 8-byte member array */

int qf_abi_version(void);
const char* qf_last_error(void);

/* ------------------------------------------------------------------------------------
 * Host side, no GPU needed: flatten ONE fitted tree into the HBM layout.
 *
 * Replaces the per-tree fit bookkeeping of quantile/ensemble.py:83-101 (the dense
 * (T,N) y_train_leaves_ / y_weights_ arrays) and the 64-byte AoS node records of
 * sklearn:tree/_tree.pxd:15-25 by
 *   nodes_out[n_nodes + 1]  8-byte packed nodes in level order with sibling pairs: slot 0 =
 *       root, slot 1 = padding (all zero, never reached), then level after level, left to
 *       right, the two children of every internal node side by side (right child = left
 *       child + 1; every pair starts on a 16-byte boundary because a tree occupies an even
 *       number of slots).  Rows that are close in feature space sit on neighbouring nodes
 *       of a level, which is what keeps the leaf-lookup warps on few 32-byte sectors.
 *       internal: lo32 = float bits of floor_f32(threshold)   (largest f32 <= the f64
 *                        threshold, so `x <= thr` on float32 x decides exactly like
 *                        sklearn:tree/_tree.pyx:989, which compares in float64)
 *                 hi32 = (left_child_offset << feature_bits) | feature,  bit31 = 0
 *       leaf:     lo32 = index of the leaf's first member in the global member array
 *                 hi32 = 0x80000000 | number_of_members
 *   node_id_out[n_nodes + 1] scikit-learn node id of every packed slot, -1 for the padding
 *       (so that `apply` can return scikit-learn ids, forest.py:583-604)
 *   members_out[n_inbag] one 8-byte entry per in-bag training sample, grouped by leaf, the
 *       leaves from left to right (depth-first order: lists of neighbouring leaves are
 *       neighbours in memory), sorted by rank inside a leaf:
 *                 lo32 = rank of the sample in np.argsort(y_train_)  (ensemble.py:133)
 *                 hi32 = float bits of y_weights_[t, sample] = f32(count / leaf_count_sum)
 *                        (ensemble.py:94-99)
 * Inputs
 *   children_left/right, feature, threshold: the arrays of sklearn's Tree (int64/f64)
 *   train_leaves[n_train]: tree_.apply(X_train)                      (tree.py:101)
 *   counts[n_train] or NULL: bootstrap multiplicities (ensemble.py:88-94); NULL = all 1
 *   sorter[n_train]: np.argsort(y_train_), computed ONCE on the host by the caller
 *   member_base: global index of this tree's first member entry
 * Outputs (besides the arrays): n_inbag_out, max_leaf_members_out, max_depth_out,
 *   sum_sq_leaf_members_out (sum over leaves of members^2: expected gather size).
 * Fails with QF_ERR_LIMIT when a child offset needs more than 31-feature_bits bits or a
 * member index exceeds 2^32-1.
 */
int qf_pack_tree(int64_t n_nodes,
                 const int64_t* children_left, const int64_t* children_right,
                 const int64_t* feature, const double* threshold,
                 int32_t feature_bits,
                 int64_t n_train, const int64_t* train_leaves, const int64_t* counts,
                 const int64_t* sorter,
                 uint64_t member_base,
                 uint64_t* nodes_out, int32_t* node_id_out, uint64_t* members_out,
                 int64_t* n_inbag_out, int32_t* max_leaf_members_out,
                 int32_t* max_depth_out, double* sum_sq_leaf_members_out);

/* Number of in-bag samples of one tree (size of members_out); counts NULL = n_train. */
int64_t qf_count_inbag(int64_t n_train, const int64_t* counts);

/* qf_pack_tree with train_leaves == NULL packs the STRUCTURE only (nodes_out, node_id_out,
 * max_depth_out; n_train / counts / sorter / members_out are ignored): leaves are left empty
 * (hi32 = 0x80000000) with lo32 = member_base + left-to-right ordinal of the leaf, for
 * qf_build_csr to fill on the device.  Pass member_base = (first packed slot of the tree) / 2:
 * a tree of k slots has k / 2 leaves, so the ordinals of a forest are dense and unique. */

/* ------------------------------------------------------------------------------------
 * Device side, fit time: the leaf -> member lists built on the GPU (K4).
 *
 * Replaces the same reference code as the member part of qf_pack_tree -- the O(T*L*N) loop
 * of quantile/ensemble.py:87-101 and tree_.apply(X_train) per tree (tree.py:98-101) -- and
 * produces bit for bit the same nodes / members arrays: leaf lookup of the training rows
 * with the predict path's own kernel, per-leaf counts, a scan over the leaves in left-to-right
 * order (the ordinals a structure-only qf_pack_tree left in lo32), a
 * scatter of (rank, f32(count / leaf count sum)) entries and a sort of every list by rank.
 * All "_dev" arrays live on `device`.  counts_dev is laid out [n_train][n_trees] (uint16
 * bootstrap multiplicities, ensemble.py:88-94) or NULL (every sample once).  Lists longer
 * than 4096 members fail with QF_ERR_LIMIT (use the host flattener).  Synchronous.
 */
typedef struct qf_csr_build_desc {
    int32_t device;
    int32_t n_trees;
    int32_t n_features;
    int32_t feature_bits;
    int64_t n_train;
    int64_t n_nodes;
    uint64_t* nodes_dev;            /* in: structure-only packed nodes; out: leaves filled  */
    const int64_t* tree_offsets;    /* host [T+1]                                           */
    const float* X_train_dev;       /* (n_train, n_features) float32 row-major              */
    const int32_t* rank_dev;        /* [n_train] rank of each sample in np.argsort(y_train_) */
    const uint16_t* counts_dev;     /* [n_train][n_trees] or NULL                           */
    uint64_t* members_dev;          /* out: (rank, weight) entries                          */
    int64_t members_capacity;       /* entries members_dev can hold                         */
    int64_t* inbag_out;             /* host [T]: members of each tree                       */
    int32_t* max_leaf_members_out;  /* host [T]                                             */
    double* sum_sq_out;             /* host [T]: sum over leaves of members^2               */
} qf_csr_build_desc;

int qf_build_csr(const qf_csr_build_desc* desc, int64_t* n_members_out);

/* ------------------------------------------------------------------------------------
 * Device side.
 */
typedef struct qf_forest qf_forest;    /* opaque: the forest resident in HBM */

typedef struct qf_forest_desc {
    int32_t device;             /* CUDA device ordinal                                  */
    int32_t n_trees;            /* T                                                    */
    int32_t n_features;         /* F                                                    */
    int32_t feature_bits;       /* as given to qf_pack_tree                             */
    int64_t n_train;            /* N                                                    */
    int64_t n_nodes;            /* sum of packed slots over trees (nodes + 1 each)      */
    int64_t n_members;          /* sum of in-bag members over trees                     */
    const uint64_t* nodes;      /* host [n_nodes]   packed nodes, trees back to back    */
    const int64_t* tree_offsets;/* host [T+1]       first packed node of each tree      */
    const int32_t* node_ids;    /* host [n_nodes]   sklearn id per packed slot (qf_pack_tree);
                                                     NULL: apply returns packed slots    */
    const uint64_t* members;    /* host [n_members] (rank, weight) entries              */
    const float* y_sorted;      /* host [N]  y_train_.astype(f32)[sorter] (utils.py:46,52) */
    const double* leaf_values;  /* host [n_nodes]   tree_.value per packed node, or NULL
                                                     (then qf_forest_predict_mean fails) */
    int64_t max_row_members;    /* upper bound on members one row can gather
                                   (sum over trees of max_leaf_members)                 */
    double expected_row_members;/* typical members per row (sum over trees of the mean
                                   leaf size seen by a training row); picks the kernel  */
} qf_forest_desc;

/* Copies everything the descriptor points at into HBM on desc->device.  Synchronous.
 * tree_offsets must be host memory; the other arrays may be host or device pointers
 * (e.g. buffers just received through an NCCL broadcast). */
int qf_forest_create(const qf_forest_desc* desc, qf_forest** out);
void qf_forest_destroy(qf_forest* f);

/* Bytes of HBM held by the forest. */
int64_t qf_forest_device_bytes(const qf_forest* f);

/* Device scratch needed to process `n_rows` rows in one call (the caller allocates it,
 * e.g. as a torch tensor, and may reuse it across calls on the same stream). */
int64_t qf_forest_workspace_bytes(const qf_forest* f, int64_t n_rows, int32_t n_quantiles);

/* Leaf lookup.  Replaces BaseForest.apply (forest.py:583-604) -> Tree._apply_dense
 * (sklearn:tree/_tree.pyx:954-996).  X_dev: (n_rows, ldx) float32 row-major, finite.
 * leaves_dev: (n_rows, T) int64 scikit-learn node ids. */
int qf_forest_apply(qf_forest* f, const float* X_dev, int64_t n_rows, int64_t ldx,
                    int64_t* leaves_dev, void* workspace_dev, int64_t workspace_bytes,
                    void* stream);

/* Quantile predict.  Replaces BaseForestQuantileRegressor.predict (ensemble.py:104-143)
 * + weighted_percentile (utils.py:3-81) for `n_quantiles` values of q at once.
 * q_host: percentiles in [0,100].  out_dev: (n_rows, n_quantiles) float64 holding the
 * reference's float32 results (ensemble.py:136,141). */
int qf_forest_predict_quantiles(qf_forest* f, const float* X_dev, int64_t n_rows, int64_t ldx,
                                const double* q_host, int32_t n_quantiles, double* out_dev,
                                int32_t mode, void* workspace_dev, int64_t workspace_bytes,
                                void* stream);

/* Mean predict.  Replaces ForestRegressor.predict (forest.py:1093-1131): float64 sum of
 * leaf values in tree order, divided by T.  out_dev: (n_rows,) float64. */
int qf_forest_predict_mean(qf_forest* f, const float* X_dev, int64_t n_rows, int64_t ldx,
                           double* out_dev, void* workspace_dev, int64_t workspace_bytes,
                           void* stream);

/* Mean and standard deviation of y | x: the reference's `return_std` forests
 * (skgarden/forest.py:133-178 `_return_std`, :332-362 RandomForestRegressor.predict, :516-548
 * ExtraTreesRegressor.predict): std = sqrt(max(0, mean_t(max(impurity, min_variance) + value^2)
 * - mean^2)), float64, trees in increasing order.  qf_forest_set_leaf_impurity uploads
 * tree_.impurity per PACKED node (host or device pointer, [n_nodes]); it must be called once
 * before qf_forest_predict_mean_std.  Outputs: (n_rows,) float64 each. */
int qf_forest_set_leaf_impurity(qf_forest* f, const double* impurity);
int qf_forest_predict_mean_std(qf_forest* f, const float* X_dev, int64_t n_rows, int64_t ldx,
                               double min_variance, double* mean_out_dev, double* std_out_dev,
                               void* workspace_dev, int64_t workspace_bytes, void* stream);

/* End-to-end variant of qf_forest_predict_quantiles on HOST buffers: chunks the rows,
 * copies each chunk host->device, runs the kernels and copies the result back,
 * pipelined over three internal streams (H2D, kernels, D2H) and a ring of chunk buffers.  X_host / out_host
 * may be pinned (copied from / to directly) or pageable (staged through page-locked slots inside the call, one
 * host memcpy per chunk that runs while the GPU is on the previous chunk).  Synchronous: returns when
 * out_host is complete.  chunk_rows <= 0 picks a default. */
int qf_forest_predict_quantiles_host(qf_forest* f, const float* X_host, int64_t n_rows,
                                     int64_t ldx, const double* q_host, int32_t n_quantiles,
                                     double* out_host, int32_t mode, int64_t chunk_rows);

/* Which kernels qf_forest_predict_quantiles runs for a call over n_rows rows (for bench
 * accounting and tests).  out[8] = {rows per chunk, keys per lane of the warp kernel (0 =
 * not used), threads of the first shared-memory stage (0 = not used), its member limit,
 * member limit of the 512-thread stage (0 = not used), dense fallback present (0/1), rows
 * walked in locality order (0/1), 0}. */
int qf_forest_plan(const qf_forest* f, int64_t n_rows, int32_t* out);

/* Name of the weight+percentile kernel that takes the bulk of the rows of a call over n_rows
 * rows with n_quantiles percentiles in `mode` (for bench accounting: the roofline line names the
 * kernel it measured).  May build the compact leaf lists of QF_MODE_FAST as a side effect. */
int qf_forest_quantile_kernel_name(qf_forest* f, int64_t n_rows, int32_t n_quantiles, int32_t mode,
                                   char* buf, int32_t len);

/* Per-call counters of the most recent predict/apply call on this handle (for bench
 * accounting): number of kernel launches issued and rows that overflowed the
 * in-shared-memory path into the dense fallback (valid after the stream is synced). */
int qf_forest_last_launches(const qf_forest* f);

/* With option "profile" = 1 the library brackets every kernel launch with CUDA events on
 * the launching stream; this returns the summed device time (ms) of the leaf-lookup
 * kernels and of the weight+percentile kernels of the most recent call.  Synchronises
 * on those events. */
int qf_forest_last_timings(qf_forest* f, double* apply_ms, double* quantile_ms);

/* Tuning knobs (bench / tests): key in {"sort_rows", "apply_row_warps", "apply_ilp",
 * "apply_levels", "rows_per_chunk", "cta_capacity", "cta2_capacity", "warp_elems", "warp_unit",
 * "select_bits", "select_impl" (0 auto, 1 multi-level, 2 select2, 3 select3, 4 select4),
 * "select3_capw", "select3_variant", "select4_hits", "select4_list", "select4_ctas",
 * "select4_unroll", "select4_hints", "compact_align", "key_bins", "dense_ctas", "profile"}.
 * Returns QF_ERR_INVALID for unknown keys or values. */
int qf_forest_set_option(qf_forest* f, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* QFOREST_B200_H */
