/*
 * ganrev_oracle.c -- CPU restatement of gan-reverser's apply_r hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see ganrev_oracle.h).  PARITY UNPINNED: no golden
 * vector of the reference exists and Torch7 cannot run here; this restates the
 * published behaviour of the Torch7 rocks at the reference's call sites and is
 * cross-checked against PyTorch-CPU / numpy (tests/golden/make_golden.py).
 *
 * Build: see oracle/Makefile (-O3 -march=x86-64-v3 -fopenmp -ffp-contract=off).
 * The file is compiled with FP contraction OFF so that every fused multiply-add
 * in the exact-match paths is an explicit fmaf(); the tolerance-graded
 * convolution loops opt back in with an attribute.
 *
 * Reference citations are file:line into aleju/gan-reverser.
 */
#include "ganrev_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BN_EPS 1e-5f /* nn.BatchNormalization default eps [upstream torch/nn] */
#define FASTFP __attribute__((optimize("fp-contract=fast")))

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ */
/* Weight blob walking (include/ganrev.h "Weight blob")                */
/* ------------------------------------------------------------------ */
typedef struct { const float *w, *b; } affine_t;               /* conv / linear: weight then bias */
typedef struct { const float *g, *b, *m, *v; } bn_t;           /* gamma, beta, running_mean, running_var */

static const float* take(const float** p, size_t n) { const float* r = *p; *p += n; return r; }
static affine_t take_affine(const float** p, size_t nw, size_t nb) {
    affine_t a; a.w = take(p, nw); a.b = take(p, nb); return a;
}
static bn_t take_bn(const float** p, size_t c) {
    bn_t b; b.g = take(p, c); b.b = take(p, c); b.m = take(p, c); b.v = take(p, c); return b;
}

size_t orc_blob_floats_G(int C, int H, int W, int nd) {
    size_t F = (size_t)512 * (H / 4) * (W / 4);
    return F * nd + F + 4 * F                         /* Linear + BN1d          models.lua:115-116 */
         + (size_t)256 * 512 * 9 + 256 + 4 * 256      /* conv1 + SBN            models.lua:122-123 */
         + (size_t)128 * 256 * 9 + 128 + 4 * 128      /* conv2 + SBN            models.lua:128-129 */
         + (size_t)C * 128 * 9 + C;                   /* conv3                  models.lua:132     */
}
size_t orc_blob_floats_R(int C, int H, int W, int nd) {
    size_t F = (size_t)128 * (H / 4) * (W / 4);
    return (size_t)64 * C * 9 + 64 + 4 * 64           /* models.lua:409-410 */
         + 2 * ((size_t)64 * 64 * 9 + 64 + 4 * 64)    /* models.lua:414-420 */
         + (size_t)128 * 64 * 9 + 128 + 4 * 128       /* models.lua:426-427 */
         + 2 * ((size_t)128 * 128 * 9 + 128 + 4 * 128)/* models.lua:431-437 */
         + 512 * F + 512 + 4 * 512                    /* models.lua:447-448 */
         + (size_t)nd * 512 + nd;                     /* models.lua:451     */
}

/* ------------------------------------------------------------------ */
/* Layers (single image, CHW)                                          */
/* ------------------------------------------------------------------ */

/* nn.SpatialConvolution / cudnn.SpatialConvolution 3x3, stride 1, pad 1. */
/* Register-blocked direct form: 4 output channels share every loaded input row and one output row stays in
 * registers across all Cin x 9 taps (the previous shift-and-accumulate form re-read and re-wrote the output plane
 * for every tap and ran at ~10 % of the cores' FMA rate).  Per output element the terms are still added in the
 * order ci ascending, ky, kx -- the same order as before, so results are unchanged up to FMA contraction. */
FASTFP static void conv3x3(const float* in, int Cin, int H, int W,
                           const float* w, const float* b, int Cout, float* out) {
    const int Wp = W + 2;
    float* pad = (float*)calloc((size_t)Cin * (H + 2) * Wp, sizeof(float));   /* zero border = the conv's padding */
    if (!pad) abort();
    for (int ci = 0; ci < Cin; ++ci)
        for (int y = 0; y < H; ++y)
            memcpy(pad + ((size_t)ci * (H + 2) + y + 1) * Wp + 1, in + ((size_t)ci * H + y) * W, sizeof(float) * W);
    enum { CB = 4, XB = 64 };
    for (int co0 = 0; co0 < Cout; co0 += CB) {
        const int nc = Cout - co0 < CB ? Cout - co0 : CB;
        for (int y = 0; y < H; ++y) {
            for (int x0 = 0; x0 < W; x0 += XB) {
                const int nx = W - x0 < XB ? W - x0 : XB;
                float acc[CB][XB];
                for (int c = 0; c < CB; ++c)
                    for (int x = 0; x < nx; ++x) acc[c][x] = c < nc ? b[co0 + c] : 0.0f;
                for (int ci = 0; ci < Cin; ++ci) {
                    for (int ky = 0; ky < 3; ++ky) {
                        const float* r = pad + ((size_t)ci * (H + 2) + y + ky) * Wp + x0;
                        for (int kx = 0; kx < 3; ++kx) {
                            float wv[CB];
                            for (int c = 0; c < CB; ++c) wv[c] = c < nc ? w[((size_t)(co0 + c) * Cin + ci) * 9 + ky * 3 + kx] : 0.0f;
                            for (int c = 0; c < CB; ++c)
                                for (int x = 0; x < nx; ++x) acc[c][x] += wv[c] * r[x + kx];
                        }
                    }
                }
                for (int c = 0; c < nc; ++c)
                    memcpy(out + ((size_t)(co0 + c) * H + y) * W + x0, acc[c], sizeof(float) * nx);
            }
        }
    }
    free(pad);
}

/* nn.(Spatial)BatchNormalization in evaluate() mode: running stats, affine. */
static void bn_eval(float* x, int Cn, int hw, bn_t bn) {
    for (int c = 0; c < Cn; ++c) {
        const float invstd = 1.0f / sqrtf(bn.v[c] + BN_EPS);
        float* p = x + (size_t)c * hw;
        for (int i = 0; i < hw; ++i) p[i] = (p[i] - bn.m[c]) * invstd * bn.g[c] + bn.b[c];
    }
}
static void relu(float* x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = x[i] > 0.0f ? x[i] : 0.0f; }
/* nn.ELU(alpha=1) */
static void elu(float* x, size_t n) { for (size_t i = 0; i < n; ++i) x[i] = x[i] > 0.0f ? x[i] : expm1f(x[i]); }
static void scale(float* x, size_t n, float s) { for (size_t i = 0; i < n; ++i) x[i] *= s; }

/* nn.SpatialUpSamplingNearest(2) */
static void upsample2(const float* in, int Cn, int H, int W, float* out) {
    const int H2 = 2 * H, W2 = 2 * W;
    for (int c = 0; c < Cn; ++c)
        for (int y = 0; y < H2; ++y)
            for (int x = 0; x < W2; ++x)
                out[((size_t)c * H2 + y) * W2 + x] = in[((size_t)c * H + (y >> 1)) * W + (x >> 1)];
}
/* nn.SpatialMaxPooling(2,2) */
static void maxpool2(const float* in, int Cn, int H, int W, float* out) {
    const int Ho = H / 2, Wo = W / 2;
    for (int c = 0; c < Cn; ++c)
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x) {
                const float* p = in + ((size_t)c * H + 2 * y) * W + 2 * x;
                float m = p[0];
                if (p[1] > m) m = p[1];
                if (p[W] > m) m = p[W];
                if (p[W + 1] > m) m = p[W + 1];
                out[((size_t)c * Ho + y) * Wo + x] = m;
            }
}
/* nn.Linear */
FASTFP static void linear(const float* x, int in, const float* w, const float* b, int out, float* y) {
    for (int o = 0; o < out; ++o) {
        const float* wr = w + (size_t)o * in;
        float acc = 0.0f;
        for (int i = 0; i < in; ++i) acc += wr[i] * x[i];
        y[o] = acc + b[o];
    }
}

/* ------------------------------------------------------------------ */
/* G: models.lua:104-143                                               */
/* ------------------------------------------------------------------ */
int orc_forward_G(const float* blob, int C, int H, int W, int nd,
                  const float* noise, int64_t N, float* images) {
    if (H % 4 || W % 4 || C < 1 || nd < 1) return 1;
    const int sH = H / 4, sW = W / 4, F = 512 * sH * sW;
    const float* p = blob;
    const affine_t lin = take_affine(&p, (size_t)F * nd, F);   /* :115 */
    const bn_t bn0 = take_bn(&p, F);                           /* :116 */
    const affine_t c1 = take_affine(&p, (size_t)256 * 512 * 9, 256); /* :122 */
    const bn_t bn1 = take_bn(&p, 256);                         /* :123 */
    const affine_t c2 = take_affine(&p, (size_t)128 * 256 * 9, 128); /* :128 */
    const bn_t bn2 = take_bn(&p, 128);                         /* :129 */
    const affine_t c3 = take_affine(&p, (size_t)C * 128 * 9, C);     /* :132 */
    const int H2 = 2 * sH, W2 = 2 * sW;
    int err = 0;
#pragma omp parallel
    {
        float* a0 = (float*)malloc(sizeof(float) * (size_t)F);
        float* u1 = (float*)malloc(sizeof(float) * (size_t)512 * H2 * W2);
        float* a1 = (float*)malloc(sizeof(float) * (size_t)256 * H2 * W2);
        float* u2 = (float*)malloc(sizeof(float) * (size_t)256 * H * W);
        float* a2 = (float*)malloc(sizeof(float) * (size_t)128 * H * W);
        if (!a0 || !u1 || !a1 || !u2 || !a2) {
#pragma omp atomic write
            err = 2;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t n = 0; n < N; ++n) {
                linear(noise + n * nd, nd, lin.w, lin.b, F, a0);     /* :115 */
                bn_eval(a0, F, 1, bn0);                              /* :116 BatchNormalization over features */
                relu(a0, F);                                         /* :117 */
                /* :118 View(512,sH,sW): feature f = c*(sH*sW) + y*sW + x -- a0 already is CHW */
                upsample2(a0, 512, sH, sW, u1);                      /* :121 */
                conv3x3(u1, 512, H2, W2, c1.w, c1.b, 256, a1);       /* :122 */
                bn_eval(a1, 256, H2 * W2, bn1);                      /* :123 */
                relu(a1, (size_t)256 * H2 * W2);                     /* :124 */
                upsample2(a1, 256, H2, W2, u2);                      /* :127 */
                conv3x3(u2, 256, H, W, c2.w, c2.b, 128, a2);         /* :128 */
                bn_eval(a2, 128, H * W, bn2);                        /* :129 */
                relu(a2, (size_t)128 * H * W);                       /* :130 */
                float* img = images + n * (int64_t)C * H * W;
                conv3x3(a2, 128, H, W, c3.w, c3.b, C, img);          /* :132 */
                for (int i = 0; i < C * H * W; ++i)                  /* :133 Sigmoid */
                    img[i] = 1.0f / (1.0f + expf(-img[i]));
            }
        }
        free(a0); free(u1); free(a1); free(u2); free(a2);
    }
    return err;
}

/* ------------------------------------------------------------------ */
/* R: models.lua:389-464                                               */
/* ------------------------------------------------------------------ */
int orc_forward_R(const float* blob, int C, int H, int W, int nd, int tanh_out,
                  const float* images, const uint8_t* mask, int64_t N, float* attrs) {
    if (H % 4 || W % 4 || C < 1 || nd < 1) return 1;
    const int Hh = H / 2, Wh = W / 2, Hq = H / 4, Wq = W / 4, F = 128 * Hq * Wq;
    const float* p = blob;
    const affine_t c1 = take_affine(&p, (size_t)64 * C * 9, 64);     const bn_t b1 = take_bn(&p, 64);   /* :409-410 */
    const affine_t c2 = take_affine(&p, (size_t)64 * 64 * 9, 64);    const bn_t b2 = take_bn(&p, 64);   /* :414-415 */
    const affine_t c3 = take_affine(&p, (size_t)64 * 64 * 9, 64);    const bn_t b3 = take_bn(&p, 64);   /* :419-420 */
    const affine_t c4 = take_affine(&p, (size_t)128 * 64 * 9, 128);  const bn_t b4 = take_bn(&p, 128);  /* :426-427 */
    const affine_t c5 = take_affine(&p, (size_t)128 * 128 * 9, 128); const bn_t b5 = take_bn(&p, 128);  /* :431-432 */
    const affine_t c6 = take_affine(&p, (size_t)128 * 128 * 9, 128); const bn_t b6 = take_bn(&p, 128);  /* :436-437 */
    const affine_t l1 = take_affine(&p, (size_t)512 * F, 512);       const bn_t b7 = take_bn(&p, 512);  /* :447-448 */
    const affine_t l2 = take_affine(&p, (size_t)nd * 512, nd);                                          /* :451 */
    int err = 0;
#pragma omp parallel
    {
        const size_t big = (size_t)64 * H * W; /* == 128*Hh*Wh*2 >= every other activation */
        float* x0 = (float*)malloc(sizeof(float) * (size_t)C * H * W);
        float* t0 = (float*)malloc(sizeof(float) * big);
        float* t1 = (float*)malloc(sizeof(float) * big);
        float* h = (float*)malloc(sizeof(float) * 512);
        if (!x0 || !t0 || !t1 || !h) {
#pragma omp atomic write
            err = 2;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int64_t n = 0; n < N; ++n) {
                const float* img = images + n * (int64_t)C * H * W;
                /* :399-406 fixer input nn.Dropout(0.5, true) forced to training: v1 dropout,
                 * x * Bernoulli mask with NO 1/(1-p) rescale [upstream]; mask is explicit here. */
                if (mask) {
                    const uint8_t* mk = mask + n * (int64_t)C * H * W;
                    for (int i = 0; i < C * H * W; ++i) x0[i] = mk[i] ? img[i] : 0.0f;
                } else {
                    memcpy(x0, img, sizeof(float) * (size_t)C * H * W);
                }
                conv3x3(x0, C, H, W, c1.w, c1.b, 64, t0);   bn_eval(t0, 64, H * W, b1);  elu(t0, (size_t)64 * H * W);   /* :409-411; :412 Dropout = identity in eval */
                conv3x3(t0, 64, H, W, c2.w, c2.b, 64, t1);  bn_eval(t1, 64, H * W, b2);  elu(t1, (size_t)64 * H * W);   /* :414-417 */
                conv3x3(t1, 64, H, W, c3.w, c3.b, 64, t0);  bn_eval(t0, 64, H * W, b3);  elu(t0, (size_t)64 * H * W);   /* :419-421 */
                maxpool2(t0, 64, H, W, t1);                                                                            /* :422 */
                conv3x3(t1, 64, Hh, Wh, c4.w, c4.b, 128, t0);  bn_eval(t0, 128, Hh * Wh, b4); elu(t0, (size_t)128 * Hh * Wh); /* :426-429 */
                conv3x3(t0, 128, Hh, Wh, c5.w, c5.b, 128, t1); bn_eval(t1, 128, Hh * Wh, b5); elu(t1, (size_t)128 * Hh * Wh); /* :431-434 */
                conv3x3(t1, 128, Hh, Wh, c6.w, c6.b, 128, t0); bn_eval(t0, 128, Hh * Wh, b6); elu(t0, (size_t)128 * Hh * Wh); /* :436-438 */
                scale(t0, (size_t)128 * Hh * Wh, 0.75f);   /* :439 SpatialDropout(0.25) in eval: x*(1-p) [upstream] */
                maxpool2(t0, 128, Hh, Wh, t1);             /* :440 */
                /* :446 View(128*Hq*Wq): CHW flatten -- t1 already is */
                linear(t1, F, l1.w, l1.b, 512, h);  bn_eval(h, 512, 1, b7);  elu(h, 512);   /* :447-450 */
                float* out = attrs + n * nd;
                linear(h, 512, l2.w, l2.b, nd, out);                                        /* :451 */
                if (tanh_out) for (int i = 0; i < nd; ++i) out[i] = tanhf(out[i]);          /* :452-454 */
            }
        }
        free(x0); free(t0); free(t1); free(h);
    }
    return err;
}

/* ------------------------------------------------------------------ */
/* Exact-match arithmetic (canonical definitions)                      */
/* ------------------------------------------------------------------ */
static inline float dot_fma(const float* a, const float* b, int d) {
    float acc = 0.0f;
    for (int i = 0; i < d; ++i) acc = fmaf(a[i], b[i], acc);
    return acc;
}
static inline float rnorm(const float* a, int d) { /* 1/(|a|^2 + 1e-12) */
    return 1.0f / (dot_fma(a, a, d) + 1e-12f);
}
static inline float cos_from(float dot, float ra, float rb) { return dot * sqrtf(ra * rb); }

/* apply_r.lua:396-400 */
float orc_cosine(const float* a, const float* b, int d) {
    return cos_from(dot_fma(a, b, d), rnorm(a, d), rnorm(b, d));
}

/* total order for "score descending, NaN last, -0 == +0, lowest id first":
 * larger key = better. */
static inline uint64_t rank_key(float s, uint32_t id) {
    uint32_t k;
    if (s != s) {
        k = 0u;
    } else {
        s = s + 0.0f; /* -0 -> +0 */
        uint32_t b; memcpy(&b, &s, 4);
        k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
    return ((uint64_t)k << 32) | (uint64_t)(0xFFFFFFFFu - id);
}
static int cmp_u64_desc(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}

/* apply_r.lua:265-282 */
int orc_search_cosine(const float* db, int64_t N, int d, const float* queries, int Q, int k,
                      int64_t* ids, float* scores) {
    if (N < 0 || N > 0xFFFFFFFFll || d < 1 || Q < 0 || k < 0) return 1;
    const int64_t kk = k < N ? k : N;
    float* rdb = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
    if (!rdb) return 2;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < N; ++j) rdb[j] = rnorm(db + j * d, d);
    int err = 0;
#pragma omp parallel
    {
        uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
        float* sc = (float*)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1));
        if (!keys || !sc) {
#pragma omp atomic write
            err = 2;
        } else {
#pragma omp for schedule(dynamic, 1)
            for (int q = 0; q < Q; ++q) {
                const float* qv = queries + (size_t)q * d;
                const float rq = rnorm(qv, d);
                for (int64_t j = 0; j < N; ++j) {
                    sc[j] = cos_from(dot_fma(qv, db + j * d, d), rq, rdb[j]);
                    keys[j] = rank_key(sc[j], (uint32_t)j);
                }
                qsort(keys, (size_t)N, sizeof(uint64_t), cmp_u64_desc);
                for (int64_t r = 0; r < k; ++r) {
                    if (r < kk) {
                        const int64_t id = (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[r] & 0xFFFFFFFFu));
                        ids[(size_t)q * k + r] = id;
                        scores[(size_t)q * k + r] = sc[id];
                    } else {
                        ids[(size_t)q * k + r] = -1;
                        scores[(size_t)q * k + r] = 0.0f;
                    }
                }
            }
        }
        free(keys); free(sc);
    }
    free(rdb);
    return err;
}

/* ------------------------------------------------------------------ */
/* kmeans: unsup.kmeans [upstream koraykv/unsup], call at apply_r.lua:198 */
/* ------------------------------------------------------------------ */
int orc_kmeans_shift(const float* x, int64_t N, int d, int64_t N_total) {
    float mx = 0.0f;
    for (int64_t i = 0; i < N * d; ++i) {
        const float a = fabsf(x[i]);
        if (!(a <= mx)) mx = a; /* NaN propagates */
    }
    if (!(mx <= 3.0e38f)) return -1; /* non-finite data */
    int e = 0;
    if (mx > 0.0f) (void)frexpf(mx, &e); /* mx = m*2^e, m in [0.5,1) => |x| < 2^e */
    int n = 0;
    while (((int64_t)1 << n) < N_total) ++n;
    int s = 62 - n - e;
    if (s > 60) s = 60;
    if (s < 0) s = 0;
    return s;
}

/* TH max along a dim: strict "!(v <= best)" scan, break on NaN -> first max / first NaN. */
static inline int kmeans_label(const float* cen, const float* c2, int k, int d, const float* x) {
    int best = 0;
    float bv = 0.0f;
    for (int j = 0; j < k; ++j) {
        const float v = dot_fma(cen + (size_t)j * d, x, d) - c2[j];
        if (j == 0 || !(v <= bv)) {
            best = j; bv = v;
            if (v != v) break;
        }
    }
    return best;
}

int orc_kmeans(const float* x, int64_t N, int d, int k, int niter, const float* init, int shift,
               float* centroids, float* total_counts, int32_t* last_labels) {
    if (N < 1 || d < 1 || k < 1 || niter < 0) return 1;
    if (shift < 0) shift = orc_kmeans_shift(x, N, d, N);
    if (shift < 0) return 3;
    const double sc = ldexp(1.0, shift);
    memcpy(centroids, init, sizeof(float) * (size_t)k * d);
    int64_t* acc = (int64_t*)malloc(sizeof(int64_t) * (size_t)k * d);
    int64_t* cnt = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    int64_t* tot = (int64_t*)calloc((size_t)k, sizeof(int64_t));
    float* c2 = (float*)malloc(sizeof(float) * (size_t)k);
    int32_t* lab = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    if (!acc || !cnt || !tot || !c2 || !lab) { free(acc); free(cnt); free(tot); free(c2); free(lab); return 2; }
    for (int it = 0; it < niter; ++it) {
        for (int j = 0; j < k; ++j) c2[j] = 0.5f * dot_fma(centroids + (size_t)j * d, centroids + (size_t)j * d, d);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < N; ++i) lab[i] = kmeans_label(centroids, c2, k, d, x + i * d);
        memset(acc, 0, sizeof(int64_t) * (size_t)k * d);
        memset(cnt, 0, sizeof(int64_t) * (size_t)k);
        /* integer accumulation: associative, so any order / partition gives the same sums */
        for (int64_t i = 0; i < N; ++i) {
            int64_t* a = acc + (size_t)lab[i] * d;
            const float* xi = x + i * d;
            for (int c = 0; c < d; ++c) a[c] += llrint((double)xi[c] * sc);
            cnt[lab[i]] += 1;
        }
        for (int j = 0; j < k; ++j) {
            if (cnt[j] != 0) /* empty clusters keep their centroid */
                for (int c = 0; c < d; ++c)
                    centroids[(size_t)j * d + c] = (float)((double)acc[(size_t)j * d + c] / ((double)cnt[j] * sc));
            tot[j] += cnt[j];
        }
    }
    for (int j = 0; j < k; ++j) total_counts[j] = (float)tot[j];
    if (last_labels) {
        if (niter > 0) memcpy(last_labels, lab, sizeof(int32_t) * (size_t)N);
        else for (int64_t i = 0; i < N; ++i) last_labels[i] = -1;
    }
    free(acc); free(cnt); free(tot); free(c2); free(lab);
    return 0;
}

/* apply_r.lua:206-218 */
int orc_assign_cosine_min(const float* x, int64_t N, int d, const float* centroids, int k,
                          int32_t* cluster, float* cosv) {
    if (N < 0 || d < 1 || k < 1) return 1;
    float* rc = (float*)malloc(sizeof(float) * (size_t)k);
    if (!rc) return 2;
    for (int j = 0; j < k; ++j) rc[j] = rnorm(centroids + (size_t)j * d, d);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const float* xi = x + i * d;
        const float rx = rnorm(xi, d);
        int best = 0; float bv = 0.0f;
        for (int j = 0; j < k; ++j) {
            /* cosineSimilarity(attributes[i], centroids[j]) */
            const float v = cos_from(dot_fma(xi, centroids + (size_t)j * d, d), rx, rc[j]);
            if (j == 0 || v < bv) { best = j; bv = v; } /* "minDist == nil or dist < minDist" */
        }
        cluster[i] = best; cosv[i] = bv;
    }
    free(rc);
    return 0;
}

/* apply_r.lua:222-243 */
int orc_cluster_members(const int32_t* cluster, const float* cosv, int64_t N, int k, int m,
                        const float* images, int px,
                        int64_t* member_ids, int32_t* member_counts, float* mean_images) {
    if (N < 0 || N > 0xFFFFFFFFll || k < 1 || m < 1) return 1;
    uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(N > 0 ? N : 1));
    if (!keys) return 2;
    for (int j = 0; j < k; ++j) {
        int64_t n = 0;
        for (int64_t i = 0; i < N; ++i) if (cluster[i] == j) keys[n++] = rank_key(cosv[i], (uint32_t)i);
        qsort(keys, (size_t)n, sizeof(uint64_t), cmp_u64_desc);   /* table.sort a[2] > b[2] */
        const int keep = (int)(n < m ? n : m);
        member_counts[j] = keep;
        for (int r = 0; r < m; ++r)
            member_ids[(size_t)j * m + r] = r < keep ? (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[r] & 0xFFFFFFFFu)) : -1;
        if (images && mean_images) {
            float* face = mean_images + (size_t)j * px;
            for (int p = 0; p < px; ++p) {
                float s = 0.0f;                                   /* torch.zeros; face:add(img) in order */
                for (int r = 0; r < keep; ++r) s = s + images[(size_t)member_ids[(size_t)j * m + r] * px + p];
                face[p] = s / (float)keep;                        /* face:div(#clusterImgs); 0/0 = NaN as in Torch */
            }
        }
    }
    free(keys);
    return 0;
}

/* apply_r.lua:366, torch.dist(a,b) = sqrt(sum((a-b)^2)) [upstream TH: float pow, double sum] */
static double l2_one(const float* x, const float* y, int px) {
    double lane[32];
    for (int l = 0; l < 32; ++l) lane[l] = 0.0;
    for (int i = 0; i < px; ++i) {
        const float dd = x[i] - y[i];
        const float sq = dd * dd;
        lane[(i >> 2) & 31] += (double)sq;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        double nxt[32];
        for (int l = 0; l < 32; ++l) nxt[l] = lane[l] + lane[l ^ off];
        memcpy(lane, nxt, sizeof(lane));
    }
    return sqrt(lane[0]);
}
int orc_l2(const float* a, const float* b, int64_t N, int px, double* l2) {
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) l2[n] = l2_one(a + n * px, b + n * px, px);
    return 0;
}
int orc_l2_sequential(const float* a, const float* b, int64_t N, int px, double* l2) {
    for (int64_t n = 0; n < N; ++n) {
        const float* x = a + n * px; const float* y = b + n * px;
        double s = 0.0;
        for (int i = 0; i < px; ++i) { const float dd = x[i] - y[i]; const float sq = dd * dd; s += (double)sq; }
        l2[n] = sqrt(s);
    }
    return 0;
}

/* sample.lua:128-148 findClosestNeighboursOf: for every query image scan the whole set with
 * torch.dist and keep the first strictly smaller distance ("closestDist == nil or dist < closestDist").
 * Row 0 is always taken first, so a NaN distance at row 0 sticks (every later "dist < NaN" is false);
 * NaN distances at later rows are never taken.  Distances use orc_l2's canonical summation order. */
int orc_nearest_l2(const float* q, int Q, const float* set, int64_t N, int px, int64_t* ids, double* dist) {
    if (Q < 0 || N < 0 || px < 1) return 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < Q; ++i) {
        int64_t best = -1;
        double bd = 0.0;
        for (int64_t j = 0; j < N; ++j) {
            const double dj = l2_one(set + j * px, q + (int64_t)i * px, px);
            if (best < 0 || dj < bd) { best = j; bd = dj; }
        }
        ids[i] = best;
        dist[i] = best < 0 ? INFINITY : bd;
    }
    return 0;
}

static int cmp_f64_asc(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
/* apply_r.lua:370-378 */
int orc_anomaly_flags(const double* l2, int64_t n_calc, int64_t n_show, double quantile,
                      uint8_t* flags, double* thr_out) {
    if (n_calc < 1 || n_show < 0 || n_show > n_calc) return 1;
    const int64_t r = (int64_t)floor((double)n_calc * quantile); /* math.floor(#distancesForSort*threshold) */
    if (r < 1 || r > n_calc) return 1;                           /* Lua would index nil */
    double* s = (double*)malloc(sizeof(double) * (size_t)n_calc);
    if (!s) return 2;
    for (int64_t i = 0; i < n_calc; ++i) s[i] = 1.0 - l2[i];     /* "1 - torch.dist(...)" */
    qsort(s, (size_t)n_calc, sizeof(double), cmp_f64_asc);       /* table.sort(distancesForSort) */
    const double thr = s[r - 1];
    for (int64_t i = 0; i < n_show; ++i) flags[i] = (1.0 - l2[i]) <= thr ? 1 : 0;
    if (thr_out) *thr_out = thr;
    free(s);
    return 0;
}
