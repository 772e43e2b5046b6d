/*
 * ganrev_oracle.h -- CPU restatement of gan-reverser's apply_r hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and there only as the checker
 * or as the timed CPU baseline.  The product (libganrev_cuda.so) never links,
 * loads or calls it.
 *
 * PARITY UNPINNED: the reference (aleju/gan-reverser) ships no tests, golden
 * vectors or model files, and its arithmetic lives in un-vendored, un-pinned
 * Torch7 rocks (nn, cudnn, unsup, torch; README.md:93-97) that cannot run in
 * this image (no lua/luajit/th).  This file restates the published algorithms
 * of those rocks at the reference's call sites; it is cross-checked against
 * an independent PyTorch-CPU / numpy statement (tests/golden/make_golden.py),
 * not against Torch7 itself.
 *
 * All tensors are contiguous row-major float32, images NCHW, ids 0-based.
 */
#ifndef GANREV_ORACLE_H
#define GANREV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Number of floats in the weight blob (layout: include/ganrev.h "Weight blob"). */
size_t orc_blob_floats_G(int C, int H, int W, int nd);
size_t orc_blob_floats_R(int C, int H, int W, int nd);

/* models.lua:104-143 (create_G3), eval mode.  noise [N x nd] -> images [N x C x H x W] in [0,1]. */
int orc_forward_G(const float* blob, int C, int H, int W, int nd,
                  const float* noise, int64_t N, float* images);

/* models.lua:389-464 (create_R_default), eval mode.  mask (may be NULL) is the
 * fixer's always-on input dropout (models.lua:399-406): x * mask, no rescale. */
int orc_forward_R(const float* blob, int C, int H, int W, int nd, int tanh_out,
                  const float* images, const uint8_t* mask, int64_t N, float* attrs);

/* apply_r.lua:396-400 (nn.CosineDistance), canonical arithmetic: sequential
 * fp32 fmaf over d in index order; w = (1/(na+1e-12))*(1/(nb+1e-12)); dot*sqrt(w). */
float orc_cosine(const float* a, const float* b, int d);

/* apply_r.lua:265-318: for each query, cosine against every db row, sorted by
 * (score desc, id asc), NaN scores last; first k kept. */
int orc_search_cosine(const float* db, int64_t N, int d, const float* queries, int Q, int k,
                      int64_t* ids, float* scores);

/* Canonical fixed-point shift for the order-free kmeans accumulators. */
int orc_kmeans_shift(const float* x, int64_t N, int d, int64_t N_total);

/* unsup.kmeans as called at apply_r.lua:198 (Lloyd, argmax of c.x - 0.5|c|^2,
 * first index on ties, empty clusters keep their centroid).  init [k x d] is
 * the explicit starting centroid set.  shift < 0 => orc_kmeans_shift(x,N,d,N). */
int orc_kmeans(const float* x, int64_t N, int d, int k, int niter, const float* init, int shift,
               float* centroids, float* total_counts, int32_t* last_labels);

/* apply_r.lua:206-218: cluster with the MINIMUM cosine, strict <, lowest j on ties. */
int orc_assign_cosine_min(const float* x, int64_t N, int d, const float* centroids, int k,
                          int32_t* cluster, float* cosv);

/* apply_r.lua:222-243: per cluster keep <= m members sorted by (cos desc, id asc);
 * mean image = sequential fp32 adds in that order, then / count.
 * member_ids [k x m] (-1 padded), member_counts [k], mean_images [k x px]. */
int orc_cluster_members(const int32_t* cluster, const float* cosv, int64_t N, int k, int m,
                        const float* images, int px,
                        int64_t* member_ids, int32_t* member_counts, float* mean_images);

/* apply_r.lua:366 torch.dist(a,b): canonical lane-tree order (lane=(i/4)%32,
 * fp32 diff, fp32 square, double partial sums, xor-butterfly), sqrt in double. */
int orc_l2(const float* a, const float* b, int64_t N, int px, double* l2);
/* sample.lua:128-148: nearest set image per query by torch.dist, first strict minimum; ids -1 / dist inf when N == 0 */
int orc_nearest_l2(const float* q, int Q, const float* set, int64_t N, int px, int64_t* ids, double* dist);
/* Same quantity in TH's plain sequential order (for bounding the order effect). */
int orc_l2_sequential(const float* a, const float* b, int64_t N, int px, double* l2);

/* apply_r.lua:370-378: sims = 1 - l2; thr = ascending sims[floor(n_calc*q)] (1-based);
 * flags[i] = sims[i] <= thr for i < n_show.  Returns nonzero if floor(n_calc*q) == 0. */
int orc_anomaly_flags(const double* l2, int64_t n_calc, int64_t n_show, double quantile,
                      uint8_t* flags, double* thr_out);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
