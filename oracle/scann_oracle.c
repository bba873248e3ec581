/*
 * scann_oracle.c -- CPU restatement of ScaNN's batched query hot path
 * (tokenize -> AH LUT -> LUT16 scan -> top-N -> SOAR dedup -> exact reorder -> sort), of bf16 / float brute force,
 * and of the per-datapoint stage of index construction (database tokenization, SOAR assignment, AH encoding).
 *
 * TEST INFRASTRUCTURE ONLY -- see scann_oracle.h.  The reference as a whole cannot be built here and ships no golden
 * vectors for this path; its self-contained arithmetic can, and is (oracle/Makefile -> oracle/_ref/libscann_ref.so,
 * compiled from /root/reference).  PINNED to that compiled reference code, bit for bit (tests/test_oracle_ref.py):
 * code packing, LUT fixed-point conversion, LUT16 int16 sums and float scores, the candidate contract, bf16 helpers,
 * the float tokenization chain (many-to-many accumulation, centre and query norms), the int8 tokenization distances,
 * the f32 / bf16 / int8 reordering distances (dims >= 8 for f32), SquaredL2Norm, and of the index build the database
 * tokenization, the SOAR assignment and the noise-shaped encoder.  RESTATED ONLY (no compilable
 * reference piece: they need Highway): the raw LUT distances of AH blocks with < 8 dims and of codebook centre 15,
 * the f32 reordering distance for dims < 8, plain AH codes of blocks with < 8 dims, k-means -- pinned by the numpy / pure-Python
 * restatements of tests/test_oracle.py, tests/test_oracle_build.py and by tests/golden/.
 *
 * Build: gcc -O3 -std=gnu11 -mavx2 -mfma -ffp-contract=off -fopenmp -shared -fPIC
 * (-ffp-contract=off matters: every FMA below is an explicit fmaf(), every
 *  "mul then add" must stay two roundings.)
 *
 * Contraction rule of the restatements.  The reference's documented build (README.md:98) is clang with
 * --copt=-mavx --copt=-mfma and no -ffp-contract flag, i.e. clang's default "on": a multiply feeding an add or a
 * subtract INSIDE ONE SOURCE EXPRESSION becomes one fused instruction (`acc - a * b`, `acc += a * b`,
 * `acc + tmp * tmp`), whatever the function's own target attribute, while operations written as separate intrinsic
 * calls or operator functions (`_mm_add_ps(acc, _mm_mul_ps(a, b))`, Highway's `add - mul * x`) stay two roundings.
 * The _ref build reproduces exactly that for the pieces it compiles (oracle/Makefile), which is how the rule was
 * checked where it matters (the scalar tails of the asymmetric and symmetric one-to-many kernels).
 *
 * Paths in comments are relative to /root/reference/scann/.
 */
#define _GNU_SOURCE
#include "scann_oracle.h"

#include <float.h>
#include <immintrin.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[512];
static __thread uint64_t g_scan_bytes;
static __thread uint64_t g_band;

static int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return 3; /* INVALID_ARGUMENT */
}
const char* so_last_error(void) { return g_err; }
uint64_t so_last_scan_bytes(void) { return g_scan_bytes; }
uint64_t so_last_boundary_band(void) { return g_band; }

struct so_index {
  so_index_desc d;
  int32_t* block_dims; /* [B] */
  uint32_t* block_off; /* [B+1] */
  float* centers_t;    /* [D][L] transposed copy for the tokenizer */
  uint32_t* leaf_off;  /* [L+1] slot offsets */
  uint32_t* leaf_dp;   /* [slots] datapoint id of each slot (datapoints_by_token) */
  uint8_t* slot_codes; /* [slots][B] resolved (primary vs SOAR) code rows */
  uint8_t** packed;    /* [L] reference-layout packed codes */
  int disjoint;
  uint32_t max_leaf;
  uint32_t n_slots;
};

/* ------------------------------------------------------------------------- */
/* index                                                                      */
/* ------------------------------------------------------------------------- */

/* hashes/internal/asymmetric_hashing_impl.cc:690-737 (CreatePackedDataset):
 * 32-datapoint groups; byte (g*B + j)*16 + m = code(32g+m, j) | code(32g+16+m, j) << 4;
 * the tail group replicates the last datapoint. */
static uint8_t* pack_leaf(const uint8_t* codes, uint32_t n, uint32_t B) {
  if (n == 0) return NULL;
  uint32_t groups = (n + 31) / 32;
  uint8_t* out = (uint8_t*)aligned_alloc(64, (((size_t)groups * B * 16) + 63) / 64 * 64 + 64);
  for (uint32_t g = 0; g < groups; ++g)
    for (uint32_t j = 0; j < B; ++j)
      for (uint32_t m = 0; m < 16; ++m) {
        uint32_t i0 = 32 * g + m, i1 = 32 * g + 16 + m;
        if (i0 >= n) i0 = n - 1;
        if (i1 >= n) i1 = n - 1;
        out[((size_t)g * B + j) * 16 + m] =
            (uint8_t)(codes[(size_t)i0 * B + j] | (codes[(size_t)i1 * B + j] << 4));
      }
  return out;
}

so_index* so_index_create(const so_index_desc* desc) {
  so_index* ix = (so_index*)calloc(1, sizeof *ix);
  ix->d = *desc;
  const uint32_t L = desc->n_leaves, B = desc->n_blocks, N = desc->n, D = desc->d;
  if (B) {
    ix->block_dims = (int32_t*)malloc(sizeof(int32_t) * B);
    ix->block_off = (uint32_t*)malloc(sizeof(uint32_t) * (B + 1));
    ix->block_off[0] = 0;
    for (uint32_t b = 0; b < B; ++b) {
      ix->block_dims[b] = desc->block_dims ? desc->block_dims[b] : (int32_t)desc->dims_per_block;
      ix->block_off[b + 1] = ix->block_off[b] + (uint32_t)ix->block_dims[b];
    }
  }
  if (L) {
    ix->centers_t = (float*)malloc(sizeof(float) * (size_t)L * D);
    for (uint32_t l = 0; l < L; ++l)
      for (uint32_t k = 0; k < D; ++k) ix->centers_t[(size_t)k * L + l] = desc->centers[(size_t)l * D + k];
    /* scann_ops/cc/scann.cc:88-98 (AddTokenizationToOptions): file order. */
    const uint32_t mult = desc->soar ? 2 : 1;
    const size_t len = (size_t)N * mult;
    ix->leaf_off = (uint32_t*)calloc(L + 1, sizeof(uint32_t));
    for (size_t j = 0; j < len; ++j) {
      int32_t t = desc->tokens[j];
      if (t < 0) continue;
      if ((uint32_t)t >= L) { fail("token %d out of range", t); free(ix); return NULL; }
      ix->leaf_off[t + 1]++;
    }
    for (uint32_t l = 0; l < L; ++l) {
      if (ix->leaf_off[l + 1] > ix->max_leaf) ix->max_leaf = ix->leaf_off[l + 1];
      ix->leaf_off[l + 1] += ix->leaf_off[l];
    }
    ix->n_slots = ix->leaf_off[L];
    ix->leaf_dp = (uint32_t*)malloc(sizeof(uint32_t) * (ix->n_slots + 1));
    uint32_t* cur = (uint32_t*)malloc(sizeof(uint32_t) * L);
    memcpy(cur, ix->leaf_off, sizeof(uint32_t) * L);
    for (size_t j = 0; j < len; ++j) {
      int32_t t = desc->tokens[j];
      if (t < 0) continue;
      ix->leaf_dp[cur[t]++] = (uint32_t)(j / mult);
    }
    free(cur);
    /* tree_x_hybrid/internal/utils.cc:61-116: disjoint iff no datapoint is in two leaves */
    ix->disjoint = 1;
    if (desc->soar)
      for (uint32_t i = 0; i < N; ++i)
        if (desc->tokens[2 * (size_t)i] >= 0 && desc->tokens[2 * (size_t)i + 1] >= 0) { ix->disjoint = 0; break; }
    if (B) {
      /* tree_ah_hybrid_residual.cc:385-396: SOAR row iff tok[2i+1] == leaf */
      ix->slot_codes = (uint8_t*)malloc((size_t)ix->n_slots * B + 1);
      ix->packed = (uint8_t**)calloc(L, sizeof(uint8_t*));
      for (uint32_t l = 0; l < L; ++l) {
        for (uint32_t s = ix->leaf_off[l]; s < ix->leaf_off[l + 1]; ++s) {
          uint32_t i = ix->leaf_dp[s];
          const uint8_t* row = desc->codes + (size_t)i * B;
          if (desc->soar && desc->soar_codes && desc->tokens[2 * (size_t)i + 1] == (int32_t)l)
            row = desc->soar_codes + (size_t)i * B;
          memcpy(ix->slot_codes + (size_t)s * B, row, B);
        }
        ix->packed[l] = pack_leaf(ix->slot_codes + (size_t)ix->leaf_off[l] * B,
                                  ix->leaf_off[l + 1] - ix->leaf_off[l], B);
      }
    }
  }
  return ix;
}

void so_index_destroy(so_index* ix) {
  if (!ix) return;
  if (ix->packed)
    for (uint32_t l = 0; l < ix->d.n_leaves; ++l) free(ix->packed[l]);
  free(ix->packed); free(ix->slot_codes); free(ix->leaf_dp); free(ix->leaf_off);
  free(ix->centers_t); free(ix->block_dims); free(ix->block_off); free(ix);
}

uint32_t so_leaf_size(const so_index* ix, uint32_t leaf) { return ix->leaf_off[leaf + 1] - ix->leaf_off[leaf]; }
const uint32_t* so_leaf_datapoints(const so_index* ix, uint32_t leaf) { return ix->leaf_dp + ix->leaf_off[leaf]; }
int so_disjoint(const so_index* ix) { return ix->disjoint; }

/* ------------------------------------------------------------------------- */
/* one-to-many float kernels (exact arithmetic restatements)                  */
/* ------------------------------------------------------------------------- */

/* distance_measures/one_to_many/one_to_many_symmetric.h:373-503
 * (DenseAccumulatingDistanceMeasureOneToManyInternalAvx2, dims >= 8): 8 FMA lanes,
 * top+bottom, optional 4- and 2-wide steps, sum4 = (x0+x2)+(x1+x3), scalar tail.
 * DotProductDistanceLambdas::FmaTerm = fnmadd (:995-1001).  The scalar tail is
 * `acc - a*b` inside an AVX2+FMA target function and is assumed to contract. */
static float neg_dot_avx2_order(const float* q, const float* x, uint32_t n) {
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8)
    for (int l = 0; l < 8; ++l) a[l] = fmaf(-q[j + l], x[j + l], a[l]);
  float b[4];
  for (int l = 0; l < 4; ++l) b[l] = a[l + 4] + a[l];
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) b[l] = fmaf(-q[j + l], x[j + l], b[l]);
    j += 4;
  }
  if (j + 2 <= n) {
    b[2] = fmaf(-q[j], x[j], b[2]);
    b[3] = fmaf(-q[j + 1], x[j + 1], b[3]);
    j += 2;
  }
  float r = (b[0] + b[2]) + (b[1] + b[3]);
  if (j < n) r = fmaf(-q[j], x[j], r);
  return r;
}

/* SquaredL2DistanceLambdas::FmaTerm (:1043-1051): tmp = a-b; fmadd(tmp,tmp,acc). */
static float sql2_avx2_order(const float* q, const float* x, uint32_t n) {
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8)
    for (int l = 0; l < 8; ++l) { float t = q[j + l] - x[j + l]; a[l] = fmaf(t, t, a[l]); }
  float b[4];
  for (int l = 0; l < 4; ++l) b[l] = a[l + 4] + a[l];
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) { float t = q[j + l] - x[j + l]; b[l] = fmaf(t, t, b[l]); }
    j += 4;
  }
  if (j + 2 <= n) {
    float t2 = q[j] - x[j], t3 = q[j + 1] - x[j + 1];
    b[2] = fmaf(t2, t2, b[2]);
    b[3] = fmaf(t3, t3, b[3]);
    j += 2;
  }
  float r = (b[0] + b[2]) + (b[1] + b[3]);
  if (j < n) { float t = q[j] - x[j]; r = fmaf(t, t, r); }
  return r;
}

/* one_to_many_symmetric.h:691-800 (Highway path taken when dims < 8; the static
 * Highway target of a plain x86-64 build has 4 f32 lanes, half vector 2 lanes),
 * NegMulAdd without FMA = acc - a*b with two roundings (Highway's 128-bit NegMulAdd below AVX2 is `add - mul * x`
 * through its operator functions, which no contraction mode fuses); ReduceSum of 4 lanes = (a0+a2)+(a1+a3); at most
 * one scalar tail step because Lanes(d) <= 4 breaks -- the lambdas' scalar AccTerm, ONE expression `acc - a * b`, which
 * the reference's documented build (README.md:98: clang, --copt=-mfma, default -ffp-contract=on) fuses. */
static float neg_dot_small(const float* q, const float* x, uint32_t n) {
  float a[4] = {0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 4 <= n; j += 4)
    for (int l = 0; l < 4; ++l) { float p = q[j + l] * x[j + l]; a[l] = a[l] - p; }
  if (j + 2 <= n) {
    float p0 = q[j] * x[j], p1 = q[j + 1] * x[j + 1];
    a[0] = a[0] - p0;
    a[1] = a[1] - p1;
    j += 2;
  }
  float r = (a[0] + a[2]) + (a[1] + a[3]);
  if (j < n) r = fmaf(-q[j], x[j], r);  /* scalar AccTerm `acc - a * b` (:1013): one expression, contracts (see the header) */
  return r;
}
static float sql2_small(const float* q, const float* x, uint32_t n) {
  float a[4] = {0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 4 <= n; j += 4)
    for (int l = 0; l < 4; ++l) { float t = q[j + l] - x[j + l]; float p = t * t; a[l] = a[l] + p; }
  if (j + 2 <= n) {
    float t0 = q[j] - x[j], t1 = q[j + 1] - x[j + 1];
    float p0 = t0 * t0, p1 = t1 * t1;
    a[0] = a[0] + p0;
    a[1] = a[1] + p1;
    j += 2;
  }
  float r = (a[0] + a[2]) + (a[1] + a[3]);
  if (j < n) { float t = q[j] - x[j]; r = fmaf(t, t, r); }  /* scalar AccTerm `acc + tmp * tmp` (:1063-1066) */
  return r;
}

/* distance_measures/one_to_one/dot_product_sse4.cc:242-296 (DenseDotProductSse4, float):
 * two 4-lane accumulators, mul then add, hadd twice = (a0+a1)+(a2+a3). */
static float dot_sse4_order(const float* q, const float* x, uint32_t n) {
  float a[4] = {0, 0, 0, 0};
  uint32_t j = 0;
  if (n >= 8) {
    float a0[4], a1[4];
    for (int l = 0; l < 4; ++l) { a0[l] = q[l] * x[l]; a1[l] = q[4 + l] * x[4 + l]; }
    j = 8;
    for (; j + 8 <= n; j += 8)
      for (int l = 0; l < 4; ++l) {
        float p0 = q[j + l] * x[j + l], p1 = q[j + 4 + l] * x[j + 4 + l];
        a0[l] = a0[l] + p0;
        a1[l] = a1[l] + p1;
      }
    for (int l = 0; l < 4; ++l) a[l] = a0[l] + a1[l];
  }
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) { float p = q[j + l] * x[j + l]; a[l] = a[l] + p; }
    j += 4;
  }
  if (j + 2 <= n) {
    float p2 = q[j] * x[j], p3 = q[j + 1] * x[j + 1];
    a[0] = a[0] + 0.0f; /* lanes 0,1 get +0*0 */
    a[1] = a[1] + 0.0f;
    a[2] = a[2] + p2;
    a[3] = a[3] + p3;
    j += 2;
  }
  if (j < n) a[0] = fmaf(q[j], x[j], a[0]);  /* `accumulator[0] += aptr[0] * bptr[0]` (:291-293): contracts */
  return (a[0] + a[1]) + (a[2] + a[3]);
}
/* l2_distance_sse4.cc (DenseSquaredL2DistanceSse4, float): same shape with (a-b)^2. */
static float sql2_sse4_order(const float* q, const float* x, uint32_t n) {
  float a[4] = {0, 0, 0, 0};
  uint32_t j = 0;
  if (n >= 8) {
    float a0[4], a1[4];
    for (int l = 0; l < 4; ++l) {
      float t0 = q[l] - x[l], t1 = q[4 + l] - x[4 + l];
      a0[l] = t0 * t0; a1[l] = t1 * t1;
    }
    j = 8;
    for (; j + 8 <= n; j += 8)
      for (int l = 0; l < 4; ++l) {
        float t0 = q[j + l] - x[j + l], t1 = q[j + 4 + l] - x[j + 4 + l];
        float p0 = t0 * t0, p1 = t1 * t1;
        a0[l] = a0[l] + p0;
        a1[l] = a1[l] + p1;
      }
    for (int l = 0; l < 4; ++l) a[l] = a0[l] + a1[l];
  }
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) { float t = q[j + l] - x[j + l]; float p = t * t; a[l] = a[l] + p; }
    j += 4;
  }
  if (j + 2 <= n) {
    float t2 = q[j] - x[j], t3 = q[j + 1] - x[j + 1];
    float p2 = t2 * t2, p3 = t3 * t3;
    a[2] = a[2] + p2;
    a[3] = a[3] + p3;
    j += 2;
  }
  if (j < n) { float t = q[j] - x[j]; a[0] = fmaf(t, t, a[0]); }  /* l2_distance_sse4.cc:214-216: contracts */
  return (a[0] + a[1]) + (a[2] + a[3]);
}

/* DenseDistanceOneToMany of one query vs a small dense set (the 16 centres of a block):
 * one_to_many_symmetric.h:659-689 dispatch: dims < 8 -> Highway path, else AVX2 path;
 * results 0..3*floor(n/3)-1 use the accumulating kernel, the remainder uses
 * lambdas.VectorVector = DistanceMeasure::GetDistanceDense (SSE4 one-to-one). */
static void one_to_many(int distance, const float* q, const float* rows, uint32_t nrows,
                        uint32_t stride, uint32_t n, float* out) {
  const uint32_t par_end = (nrows / 3) * 3;
  for (uint32_t i = 0; i < nrows; ++i) {
    const float* x = rows + (size_t)i * stride;
    float r;
    if (i < par_end) {
      if (distance == SO_DOT_PRODUCT) r = n < 8 ? neg_dot_small(q, x, n) : neg_dot_avx2_order(q, x, n);
      else r = n < 8 ? sql2_small(q, x, n) : sql2_avx2_order(q, x, n);
    } else {
      if (distance == SO_DOT_PRODUCT) r = (float)(-(double)dot_sse4_order(q, x, n));
      else r = sql2_sse4_order(q, x, n);
    }
    out[i] = r;
  }
}

/* ------------------------------------------------------------------------- */
/* tokenization                                                               */
/* ------------------------------------------------------------------------- */

/* distance_measures/many_to_many/many_to_many_impl.inc:522-567
 * (DoAccumulationTransposedTemplate): dot: acc = 0; for dim: acc = fnmadd(q[dim], c[dim], acc).
 * squared L2: acc = ||c||^2 + ||q||^2 with ||c||^2 = -(fnmadd chain) (:236-257) and
 * acc = fnmadd(q[dim], 2*c[dim], acc).  Sequential in dim. */
static float neg_dot_i8_order(const float* qp, const int8_t* x, uint32_t n);
static double squared_l2_norm_f64(const float* v, uint32_t n);
/* One-to-one int8 x float dot product, DenseDotProductInt8FloatAvxImpl<AvxFunctionsAvx2Fma>
 * (distance_measures/one_to_one/dot_product_impl.inc:3-54): two 8-lane fmadd accumulators over whole groups of 16 dims
 * (lanes 0-7 -> acc0, 8-15 -> acc1), one 8-wide fmadd step into acc0, one 4-wide step as a rounded product ADDED into
 * lanes 0..3 of acc0 (mul_ps + add_ps, not fused), Sum8(acc0 + acc1) = ((x0+x4)+(x2+x6)) + ((x1+x5)+(x3+x7))
 * (utils/internal/avx2_funcs.h:56-63), the last < 4 dims as `acc += float(a) * b` (inside an AVX2+FMA target function:
 * assumed to contract, as the other scalar tails of this file). */
static float dot_i8_one_to_one(const float* qp, const int8_t* x, uint32_t n) {
  float a0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 16 <= n; j += 16)
    for (int l = 0; l < 8; ++l) {
      a0[l] = fmaf((float)x[j + l], qp[j + l], a0[l]);
      a1[l] = fmaf((float)x[j + 8 + l], qp[j + 8 + l], a1[l]);
    }
  if (j + 8 <= n) {
    for (int l = 0; l < 8; ++l) a0[l] = fmaf((float)x[j + l], qp[j + l], a0[l]);
    j += 8;
  }
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) { const float p = (float)x[j + l] * qp[j + l]; a0[l] = a0[l] + p; }
    j += 4;
  }
  float v[8];
  for (int l = 0; l < 8; ++l) v[l] = a0[l] + a1[l];
  float r = ((v[0] + v[4]) + (v[2] + v[6])) + ((v[1] + v[5]) + (v[3] + v[7]));
  for (; j < n; ++j) r = fmaf((float)x[j], qp[j], r);
  return r;
}

/* Int8 centres (query_tokenization_type FIXED_POINT_INT8).  A non-FLOAT tokenization type leaves the batched path
 * (KMeansTreePartitioner::SupportsLowLevelQueryBatching, partitioning/kmeans_tree_partitioner.h:230-236): every query
 * goes through KMeansTree::Tokenize -> KMeansTreeNode::FindChildrenWithSpilling<float, int8_t> ->
 * GetAllDistancesInt8 (trees/kmeans_tree/kmeans_tree_node.h:222-256): the query is scaled by the inverse multipliers
 * (times 2 for squared L2: `q *= inv_mult * 2`), distances = DenseDotProductDistanceOneToManyInt8Float over ALL
 * centres = OneToManyAsymmetricTemplate<dims, 3, no indices, dot product, int8_t>
 * (distance_measures/one_to_many/one_to_many_asymmetric_impl.inc:531-671): centres [0, 3 * (L / 3)) through the
 * three-at-a-time kernel (-<q', float(c)> in the order of neg_dot_i8_order, the kernel of the int8 reordering), the
 * last L mod 3 centres through ComputeOneToOneScore<0, false> = -DenseDotProductAvx2(int8, float) (:629-639,223-258),
 * which sums in another order.  Squared L2 adds (SquaredL2Norm(q) + SquaredL2Norm(float centre)) to each.  The top-P
 * selection (PostprocessDistancesForSpilling, kmeans_tree_node.cc:128-159: FastTopNeighbors(max_centers), PushBlock)
 * is the exact (distance, index) selection of the float path. */
static void center_distances_i8(const so_index* ix, const float* q, float* out) {
  const uint32_t L = ix->d.n_leaves, D = ix->d.d;
  float qp[D];
  const int l2 = ix->d.distance != SO_DOT_PRODUCT;
  for (uint32_t j = 0; j < D; ++j) {
    const float inv = ix->d.centers_inv_mult[j];
    qp[j] = l2 ? q[j] * (inv * 2.0f) : q[j] * inv;
  }
  const float qn = l2 ? (float)squared_l2_norm_f64(q, D) : 0.0f;
  const uint32_t L3 = L / 3 * 3;
  for (uint32_t l = 0; l < L; ++l) {
    const int8_t* c = ix->d.centers_i8 + (size_t)l * D;
    const float val = l < L3 ? neg_dot_i8_order(qp, c, D) : -dot_i8_one_to_one(qp, c, D);
    out[l] = l2 ? val + (qn + ix->d.centers_sqnorm[l]) : val;
  }
}

/* KMeansTreeNode::CreateFixedPointCenters (trees/kmeans_tree/kmeans_tree_node.cc:267-281):
 * ScalarQuantizeFloatDataset(float_centers, 1.0, NaN) (utils/scalar_quantization_helpers.cc:39-63,94-145):
 * multiplier[d] = 127 / max_l |c[l][d]| (1 for an all-zero column), value = Int8Quantize(c * multiplier) =
 * clamp(std::round(.), -128, 127) (scalar_quantization_helpers.h:40-50), inverse multiplier = 1.0f / multiplier;
 * sqnorm[l] = float(SquaredL2Norm(float centre)). */
int so_quantize_centers(const float* centers, uint32_t L, uint32_t D, int8_t* out_i8, float* out_inv_mult,
                        float* out_sqnorm) {
  float* mult = (float*)malloc(sizeof(float) * (D ? D : 1));
  if (!mult) return fail("out of memory");
  for (uint32_t j = 0; j < D; ++j) mult[j] = 0.0f;
  for (uint32_t l = 0; l < L; ++l)
    for (uint32_t j = 0; j < D; ++j) {
      const float a = fabsf(centers[(size_t)l * D + j]);
      if (a > mult[j]) mult[j] = a;
    }
  for (uint32_t j = 0; j < D; ++j) mult[j] = mult[j] == 0.0f ? 1.0f : 127.0f / mult[j];
  for (uint32_t l = 0; l < L; ++l) {
    for (uint32_t j = 0; j < D; ++j) {
      const float r = roundf(centers[(size_t)l * D + j] * mult[j]);
      out_i8[(size_t)l * D + j] = (int8_t)(r > 127.0f ? 127.0f : (r < -128.0f ? -128.0f : r));
    }
    out_sqnorm[l] = (float)squared_l2_norm_f64(centers + (size_t)l * D, D);
  }
  for (uint32_t j = 0; j < D; ++j) out_inv_mult[j] = 1.0f / mult[j];
  free(mult);
  return 0;
}

static void center_distances(const so_index* ix, const float* q, float* out) {
  const uint32_t L = ix->d.n_leaves, D = ix->d.d;
  if (ix->d.centers_i8) { center_distances_i8(ix, q, out); return; }
  if (ix->d.distance == SO_DOT_PRODUCT) {
    for (uint32_t l = 0; l < L; ++l) out[l] = 0.0f;
    for (uint32_t k = 0; k < D; ++k) {
      const float nq = -q[k];
      const float* c = ix->centers_t + (size_t)k * L;
      for (uint32_t l = 0; l < L; ++l) out[l] = fmaf(nq, c[l], out[l]);
    }
  } else {
    /* query_norms[i] = SquaredL2Norm(query i) (many_to_many_impl.inc:417-426; distance_measures/one_to_one/
     * l2_distance.h:108-120 -> DenseSingleAccumulate, utils/reduction.h:357-390: four strided double accumulators),
     * narrowed to float. */
    const float qnf = (float)squared_l2_norm_f64(q, D);
    for (uint32_t l = 0; l < L; ++l) out[l] = 0.0f;
    for (uint32_t k = 0; k < D; ++k) {
      const float* c = ix->centers_t + (size_t)k * L;
      for (uint32_t l = 0; l < L; ++l) out[l] = fmaf(-c[l], c[l], out[l]);
    }
    for (uint32_t l = 0; l < L; ++l) out[l] = (out[l] * -1.0f) + qnf;
    for (uint32_t k = 0; k < D; ++k) {
      const float nq = -q[k];
      const float* c = ix->centers_t + (size_t)k * L;
      for (uint32_t l = 0; l < L; ++l) { float c2 = c[l] * 2.0f; out[l] = fmaf(nq, c2, out[l]); }
    }
  }
}

/* float -> u32 that sorts like the float under DistanceComparator
 * (utils/util_functions.h:94-107): -0.0 and +0.0 compare equal, so canonicalise. */
static inline uint32_t f2ord(float f) {
  f = f + 0.0f;
  uint32_t u;
  memcpy(&u, &f, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return x < y ? -1 : x > y;
}

/* keep the `keep` smallest keys of v[0..n) in v[0..keep), unordered. */
static void select_smallest(uint64_t* v, size_t n, size_t keep) {
  if (keep >= n) return;
  size_t lo = 0, hi = n; /* invariant: answer boundary `keep` lies in [lo, hi) */
  while (hi - lo > 16) {
    uint64_t a = v[lo], b = v[lo + (hi - lo) / 2], c = v[hi - 1];
    uint64_t piv = a < b ? (b < c ? b : (a < c ? c : a)) : (a < c ? a : (b < c ? c : b));
    size_t i = lo, j = hi - 1;
    for (;;) {
      while (v[i] < piv) ++i;
      while (v[j] > piv) --j;
      if (i >= j) break;
      uint64_t t = v[i]; v[i] = v[j]; v[j] = t;
      ++i; --j;
    }
    /* [lo, j] <= piv, [j+1, hi) >= piv */
    const size_t plo = lo, phi = hi;
    if (keep <= j + 1) hi = j + 1; else lo = j + 1;
    if (hi - lo <= 1) return;
    if (lo == plo && hi == phi) break; /* no progress (degenerate pivot): sort the range */
  }
  qsort(v + lo, hi - lo, sizeof(uint64_t), cmp_u64);
}

/* partitioning/kmeans_tree_partitioner.cc:701-730 (FIXED_NUMBER_OF_CENTERS): the
 * max_centers nearest centres under FastTopNeighbors semantics = exact top-P with ties
 * broken by the smaller centre index (utils/fast_top_neighbors_impl.inc:345-374).
 * We return them sorted by (distance, leaf); the reference leaves them unsorted. */
static uint32_t tokenize_one(const so_index* ix, const float* q, int P, float* dist_scratch,
                             uint64_t* key_scratch, int32_t* out_leaf, float* out_dist) {
  const uint32_t L = ix->d.n_leaves;
  center_distances(ix, q, dist_scratch);
  for (uint32_t l = 0; l < L; ++l) key_scratch[l] = ((uint64_t)f2ord(dist_scratch[l]) << 32) | l;
  uint32_t p = (uint32_t)P < L ? (uint32_t)P : L;
  select_smallest(key_scratch, L, p);
  qsort(key_scratch, p, sizeof(uint64_t), cmp_u64);
  for (uint32_t i = 0; i < p; ++i) {
    out_leaf[i] = (int32_t)(key_scratch[i] & 0xffffffffu);
    out_dist[i] = dist_scratch[out_leaf[i]];
  }
  return p;
}

int so_tokenize(const so_index* ix, const float* q, uint32_t nq, int leaves, int32_t* out_leaf,
                float* out_dist) {
  const uint32_t L = ix->d.n_leaves, D = ix->d.d;
  int P = leaves > 0 ? leaves : ix->d.default_leaves;
  if ((uint32_t)P > L) P = (int)L;
  float* ds = (float*)malloc(sizeof(float) * L);
  uint64_t* ks = (uint64_t*)malloc(sizeof(uint64_t) * L);
  for (uint32_t i = 0; i < nq; ++i)
    tokenize_one(ix, q + (size_t)i * D, P, ds, ks, out_leaf + (size_t)i * P, out_dist + (size_t)i * P);
  free(ds); free(ks);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* lookup table                                                               */
/* ------------------------------------------------------------------------- */

/* hashes/internal/asymmetric_hashing_impl.cc:505-569 (CreateRawFloatLookupTable, generic
 * one-to-many path because 16-centre models never build block-transposed centres,
 * hashes/asymmetric_hashing2/training_model.cc:157-159), :572-587 (multiplier, quantile 1.0),
 * :589-645 (ConvertLookupToFixedPoint<uint8_t>, ROUND): lut = u8(round(raw*mult) + 128). */
void so_lut_quantize(const float* raw, uint64_t n, uint8_t* lut, float* mult_out) {
  float maxabs = 0.0f;
  for (uint64_t i = 0; i < n; ++i) { float a = fabsf(raw[i]); if (a > maxabs) maxabs = a; }
  const float floor_ = sqrtf(FLT_EPSILON);
  const float denom = maxabs > floor_ ? maxabs : floor_;
  const float mult = 127 / denom;
  for (uint64_t i = 0; i < n; ++i) {
    float v = raw[i] * mult;
    float r = roundf(v) + 128;
    lut[i] = (uint8_t)r;
  }
  *mult_out = mult;
}

static void lut_one(const so_index* ix, const float* q, uint8_t* lut, float* mult_out, float* raw) {
  const uint32_t B = ix->d.n_blocks, S = ix->d.dims_per_block;
  for (uint32_t b = 0; b < B; ++b)
    one_to_many(ix->d.distance, q + ix->block_off[b], ix->d.codebook + (size_t)b * 16 * S, 16, S,
                (uint32_t)ix->block_dims[b], raw + b * 16);
  so_lut_quantize(raw, (uint64_t)B * 16, lut, mult_out);
}

int so_lut(const so_index* ix, const float* q, uint32_t nq, uint8_t* out_lut, float* out_mult) {
  const uint32_t B = ix->d.n_blocks, D = ix->d.d;
  float* raw = (float*)malloc(sizeof(float) * B * 16);
  for (uint32_t i = 0; i < nq; ++i)
    lut_one(ix, q + (size_t)i * D, out_lut + (size_t)i * B * 16, out_mult + i, raw);
  free(raw);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* LUT16 scan                                                                 */
/* ------------------------------------------------------------------------- */

/* hashes/internal/lut16_avx2.inc:55-124: acc16 = sum_b lut[b][code_b] - 128*B. */
static inline int32_t score_slot(const uint8_t* lut, const uint8_t* codes, uint32_t B) {
  int32_t s = 0;
  for (uint32_t b = 0; b < B; ++b) s += lut[b * 16 + codes[b]];
  return s - 128 * (int32_t)B;
}

int so_leaf_scores(const so_index* ix, const uint8_t* lut, uint32_t leaf, int16_t* out) {
  const uint32_t B = ix->d.n_blocks;
  uint32_t n = so_leaf_size(ix, leaf);
  for (uint32_t s = 0; s < n; ++s)
    out[s] = (int16_t)score_slot(lut, ix->slot_codes + (size_t)(ix->leaf_off[leaf] + s) * B, B);
  return (int)n;
}

/* lut16_avx2.inc:429,472-476: dist = float(acc) * float(1.0 / mult) + bias (mul, then add). */
static inline float ah_float_score(int32_t acc, float inv_mult, float bias) {
  float m = (float)acc * inv_mult;
  return m + bias;
}

/* Dot-product tree-AH (TreeAHHybridResidual): inv = float(1.0 / double(mult)) (lut16_avx2.inc:429),
 * ties broken by the packed (leaf, slot) global-top-N index (tree_ah_hybrid_residual.h:234-247).
 * Squared-L2 tree-AH (TreeXHybridSMMD, non-residual codes): per-leaf int16 top-N, then
 * dist = acc * (1.0f / mult) with NO bias (hashes/asymmetric_hashing2/querying.h:450-455), pushed
 * into the per-query FastTopNeighbors<float> under GLOBAL datapoint ids
 * (base/single_machine_base.cc:759-808), so ties are broken by datapoint id. */
static inline float inv_multiplier(const so_index* ix, float mult) {
  if (ix->d.distance == SO_SQUARED_L2) return 1.0f / mult;
  return (float)(1.0 / (double)mult);
}
static inline uint32_t tie_index(const so_index* ix, uint32_t gslot) {
  return ix->d.distance == SO_SQUARED_L2 ? ix->leaf_dp[gslot] : gslot;
}

/* exact top-N collector on (score, global slot) keys: the contract of SURVEY 7.1 /
 * utils/fast_top_neighbors.h:90-118,530-548 FinishUnsorted (N smallest, ties -> smaller
 * index).  Buffer of 2N like the reference; threshold tightened on every collection. */
typedef struct {
  uint64_t* buf;
  size_t cap, n, keep;
  uint64_t thr; /* push only keys < thr */
} topn_t;
static void topn_init(topn_t* t, size_t keep) {
  t->keep = keep;
  t->cap = ((2 * keep + 31) / 32) * 32;
  if (t->cap < 64) t->cap = 64;
  t->buf = (uint64_t*)malloc(sizeof(uint64_t) * t->cap);
  t->n = 0;
  t->thr = ~(uint64_t)0;
}
static void topn_collect(topn_t* t) {
  if (t->n <= t->keep) return;
  select_smallest(t->buf, t->n, t->keep);
  t->n = t->keep;
  uint64_t mx = 0;
  for (size_t i = 0; i < t->n; ++i) if (t->buf[i] > mx) mx = t->buf[i];
  t->thr = mx;
}
static inline void topn_push(topn_t* t, uint64_t key) {
  if (key >= t->thr) return;
  t->buf[t->n++] = key;
  if (t->n == t->cap) topn_collect(t);
}
static void topn_finish(topn_t* t) {
  topn_collect(t);
  qsort(t->buf, t->n, sizeof(uint64_t), cmp_u64);
}

/* largest acc (int) whose float score is <= the score part of thr; INT32_MIN if none.
 * Conservative integer pre-filter: unlike lut16_avx2.inc:432-438 it never drops a
 * candidate that the exact key comparison would keep. */
static int32_t int_threshold(uint64_t thr, float mult, float inv_mult, float bias) {
  if (thr == ~(uint64_t)0) return 32767;
  float ts = ord2f((uint32_t)(thr >> 32));
  float est = (ts - bias) * mult;
  int32_t t;
  if (!(est < 40000.0f)) t = 32767; else if (!(est > -40000.0f)) t = -32769; else t = (int32_t)floorf(est);
  if (t > 32767) t = 32767;
  if (t < -32769) t = -32769;
  while (t < 32767 && ah_float_score(t + 1, inv_mult, bias) <= ts) ++t;
  while (t >= -32768 && ah_float_score(t, inv_mult, bias) > ts) --t;
  return t;
}

static void scan_leaf_scalar(const so_index* ix, uint32_t leaf, const uint8_t* lut, float mult,
                             float bias, topn_t* tn) {
  const uint32_t B = ix->d.n_blocks;
  const uint32_t base = ix->leaf_off[leaf], n = ix->leaf_off[leaf + 1] - base;
  const float inv_mult = inv_multiplier(ix, mult);
  for (uint32_t s = 0; s < n; ++s) {
    int32_t acc = score_slot(lut, ix->slot_codes + (size_t)(base + s) * B, B);
    float sc = ah_float_score(acc, inv_mult, bias);
    topn_push(tn, ((uint64_t)f2ord(sc) << 32) | tie_index(ix, base + s));
  }
}

/* AVX2 kernel over the reference's packed layout for up to 3 queries per pass
 * (lut16_avx2.inc:55-124 Avx2LUT16BottomLoop; tree_ah_hybrid_residual.cc:730-770 kMaxBatch=3).
 * Two blocks (32 bytes) per iteration, vpshufb per 128-bit lane, int16 accumulate with the
 * odd bytes "tagging along" in the even accumulator and subtracted at the end. */
#define SO_MAXQ 3
static inline void lut16_group_avx2(const uint8_t* data, const uint8_t* const* luts, int nqb,
                                    uint32_t B, int16_t out[SO_MAXQ][32]) {
  const __m256i low4 = _mm256_set1_epi8(0x0F);
  __m256i acc[SO_MAXQ][4];
  for (int j = 0; j < nqb; ++j) for (int k = 0; k < 4; ++k) acc[j][k] = _mm256_setzero_si256();
  uint32_t b = 0;
  for (; b + 2 <= B; b += 2) {
    __m256i codes = _mm256_loadu_si256((const __m256i*)(data + (size_t)b * 16));
    __m256i lo = _mm256_and_si256(codes, low4);
    __m256i hi = _mm256_and_si256(_mm256_srli_epi16(codes, 4), low4);
    for (int j = 0; j < nqb; ++j) {
      __m256i dict = _mm256_loadu_si256((const __m256i*)(luts[j] + (size_t)b * 16));
      __m256i r0 = _mm256_shuffle_epi8(dict, lo);
      __m256i r1 = _mm256_shuffle_epi8(dict, hi);
      acc[j][0] = _mm256_add_epi16(acc[j][0], r0);
      acc[j][1] = _mm256_add_epi16(acc[j][1], _mm256_srli_epi16(r0, 8));
      acc[j][2] = _mm256_add_epi16(acc[j][2], r1);
      acc[j][3] = _mm256_add_epi16(acc[j][3], _mm256_srli_epi16(r1, 8));
    }
  }
  for (int j = 0; j < nqb; ++j) {
    __m256i even0 = _mm256_sub_epi16(acc[j][0], _mm256_slli_epi16(acc[j][1], 8));
    __m256i even1 = _mm256_sub_epi16(acc[j][2], _mm256_slli_epi16(acc[j][3], 8));
    /* fold the two 128-bit lanes (block b and block b+1 partial sums) */
    __m128i e0 = _mm_add_epi16(_mm256_castsi256_si128(even0), _mm256_extracti128_si256(even0, 1));
    __m128i o0 = _mm_add_epi16(_mm256_castsi256_si128(acc[j][1]), _mm256_extracti128_si256(acc[j][1], 1));
    __m128i e1 = _mm_add_epi16(_mm256_castsi256_si128(even1), _mm256_extracti128_si256(even1, 1));
    __m128i o1 = _mm_add_epi16(_mm256_castsi256_si128(acc[j][3]), _mm256_extracti128_si256(acc[j][3], 1));
    if (b < B) { /* odd block count: one 16-byte tail block */
      __m128i codes = _mm_loadu_si128((const __m128i*)(data + (size_t)b * 16));
      __m128i l4 = _mm_set1_epi8(0x0F);
      __m128i lo = _mm_and_si128(codes, l4);
      __m128i hi = _mm_and_si128(_mm_srli_epi16(codes, 4), l4);
      __m128i dict = _mm_loadu_si128((const __m128i*)(luts[j] + (size_t)b * 16));
      __m128i r0 = _mm_shuffle_epi8(dict, lo);
      __m128i r1 = _mm_shuffle_epi8(dict, hi);
      __m128i m8 = _mm_set1_epi16(0x00FF);
      e0 = _mm_add_epi16(e0, _mm_and_si128(r0, m8));
      o0 = _mm_add_epi16(o0, _mm_srli_epi16(r0, 8));
      e1 = _mm_add_epi16(e1, _mm_and_si128(r1, m8));
      o1 = _mm_add_epi16(o1, _mm_srli_epi16(r1, 8));
    }
    const __m128i tb = _mm_set1_epi16((short)(128 * B));
    /* e0 holds dps 0,2,..14; o0 dps 1,3,..15; e1/o1 the same for dps 16..31 */
    __m128i d0 = _mm_sub_epi16(_mm_unpacklo_epi16(e0, o0), tb);
    __m128i d1 = _mm_sub_epi16(_mm_unpackhi_epi16(e0, o0), tb);
    __m128i d2 = _mm_sub_epi16(_mm_unpacklo_epi16(e1, o1), tb);
    __m128i d3 = _mm_sub_epi16(_mm_unpackhi_epi16(e1, o1), tb);
    _mm_storeu_si128((__m128i*)(out[j] + 0), d0);
    _mm_storeu_si128((__m128i*)(out[j] + 8), d1);
    _mm_storeu_si128((__m128i*)(out[j] + 16), d2);
    _mm_storeu_si128((__m128i*)(out[j] + 24), d3);
  }
}

static void scan_leaf_avx2(const so_index* ix, uint32_t leaf, int nqb, const uint8_t* const* luts,
                           const float* mults, const float* biases, topn_t** tns) {
  const uint32_t B = ix->d.n_blocks;
  const uint32_t base = ix->leaf_off[leaf], n = ix->leaf_off[leaf + 1] - base;
  const uint32_t groups = (n + 31) / 32;
  const uint8_t* packed = ix->packed[leaf];
  float inv[SO_MAXQ];
  int32_t thr[SO_MAXQ];
  uint64_t seen_thr[SO_MAXQ];
  for (int j = 0; j < nqb; ++j) {
    inv[j] = inv_multiplier(ix, mults[j]);
    seen_thr[j] = tns[j]->thr;
    thr[j] = int_threshold(seen_thr[j], mults[j], inv[j], biases[j]);
  }
  int16_t sc[SO_MAXQ][32] __attribute__((aligned(32)));
  for (uint32_t g = 0; g < groups; ++g) {
    lut16_group_avx2(packed + (size_t)g * B * 16, luts, nqb, B, sc);
    const uint32_t lim = (g == groups - 1) ? n - 32 * g : 32;
    for (int j = 0; j < nqb; ++j) {
      __m256i t = _mm256_set1_epi16((short)(thr[j] > 32767 ? 32767 : (thr[j] < -32768 ? -32768 : thr[j])));
      __m256i v0 = _mm256_load_si256((const __m256i*)sc[j]);
      __m256i v1 = _mm256_load_si256((const __m256i*)(sc[j] + 16));
      /* keep acc <= thr  <=>  !(acc > thr) */
      uint32_t m0 = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpgt_epi16(v0, t));
      uint32_t m1 = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpgt_epi16(v1, t));
      if (thr[j] < -32768) { m0 = 0; m1 = 0; }
      if (!(m0 | m1)) continue;
      for (uint32_t o = 0; o < lim; ++o) {
        uint32_t bit = o < 16 ? (m0 >> (2 * o)) & 1 : (m1 >> (2 * (o - 16))) & 1;
        if (!bit) continue;
        float s = ah_float_score(sc[j][o], inv[j], biases[j]);
        topn_push(tns[j], ((uint64_t)f2ord(s) << 32) | tie_index(ix, base + 32 * g + o));
        if (tns[j]->thr != seen_thr[j]) {
          seen_thr[j] = tns[j]->thr;
          thr[j] = int_threshold(seen_thr[j], mults[j], inv[j], biases[j]);
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------- */
/* whole path                                                                 */
/* ------------------------------------------------------------------------- */

typedef struct { uint32_t dp; float score; } cand_t;

static int cmp_cand_dp(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  return x->dp < y->dp ? -1 : x->dp > y->dp;
}

/* tree_x_hybrid/internal/utils.cc:135-156 (DeduplicateDatabaseSpilledResults): merge equal
 * ids with 0.5*a + 0.5*b, then keep the final_size smallest by (distance, id). */
static size_t soar_dedup(cand_t* c, size_t n, size_t final_size, uint64_t* keys) {
  qsort(c, n, sizeof(cand_t), cmp_cand_dp);
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    if (m && c[m - 1].dp == c[i].dp) {
      float a = 0.5f * c[m - 1].score, b = 0.5f * c[i].score;
      c[m - 1].score = a + b;
    } else c[m++] = c[i];
  }
  for (size_t i = 0; i < m; ++i) keys[i] = ((uint64_t)f2ord(c[i].score) << 32) | c[i].dp;
  if (m > final_size) { select_smallest(keys, m, final_size); m = final_size; }
  qsort(keys, m, sizeof(uint64_t), cmp_u64);
  for (size_t i = 0; i < m; ++i) { c[i].dp = (uint32_t)keys[i]; c[i].score = ord2f((uint32_t)(keys[i] >> 32)); }
  return m;
}

/* utils/reordering_helper.cc:257-283 -> DenseDistanceOneToMany over the gathered rows.
 * The reference's arithmetic for a row depends on its position in the (unordered) result
 * list (main kernel vs. the n%3 tail); the oracle uses the main kernel for every row. */
/* utils/bfloat16_helpers.h:30-48 (Bfloat16Decompress): bits << 16. */
static inline float bf16_to_f32(int16_t b) {
  uint32_t u = (uint32_t)(uint16_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static float neg_dot_bf16_avx2_order(const float* q, const int16_t* x, uint32_t n);
static float sql2_bf16_avx2_order(const float* q, const int16_t* x, uint32_t n);
static double squared_l2_norm_f64(const float* v, uint32_t n);

/* int8 (fixed point) reordering: FixedPointFloatDense{DotProduct,SquaredL2}ReorderingHelper::
 * ComputeDistancesForReordering (utils/reordering_helper.cc:430-441,610-618):
 *   q'[d] = (1.0f / multiplier[d]) * q[d]     (PrepareForAsymmetricScalarQuantizedDotProduct,
 *                                              utils/scalar_quantization_helpers.cc:341-351; the inverse is formed
 *                                              once in the helper's constructor, reordering_helper.cc:407-412)
 *   val = -<q', float(x)> in the order of OneToManyAsymmetricTemplate<.., int8_t> on AVX2
 *         (distance_measures/one_to_many/one_to_many_asymmetric_impl.inc:296-353): eight fnmadd lanes over whole
 *         groups of 8 dims (HandleXDims<16> is two such steps), one 4-wide step into lanes 0..3, HorizontalSum3X
 *         = ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)) (utils/intrinsics/horizontal_sum.h:170-182), then the remaining
 *         dims one by one with fnmadd on the scalar.
 *   dot product: val.  squared L2: (|q|^2 + dp_norms[id]) + 2.0f * val  (SetSquaredL2DistanceFunctor,
 *   reordering_helper.cc:108-131; |q|^2 = float(SquaredL2Norm(q))).
 * As for the float kernel, the reference sends the last n mod 3 rows of a list through a one-to-one kernel with
 * another summation order; the list order is unspecified there, so every row uses the main kernel here. */
static float neg_dot_i8_order(const float* qp, const int8_t* x, uint32_t n) {
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8)
    for (int l = 0; l < 8; ++l) a[l] = fmaf(-qp[j + l], (float)x[j + l], a[l]);
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) a[l] = fmaf(-qp[j + l], (float)x[j + l], a[l]);
    j += 4;
  }
  float r = ((a[0] + a[4]) + (a[2] + a[6])) + ((a[1] + a[5]) + (a[3] + a[7]));
  for (; j < n; ++j) r = fmaf(-qp[j], (float)x[j], r);
  return r;
}
static float exact_distance_i8(const so_index* ix, const float* q, uint32_t dp) {
  const uint32_t D = ix->d.d;
  float qp[D];
  for (uint32_t j = 0; j < D; ++j) { const float inv = 1.0f / ix->d.int8_multipliers[j]; qp[j] = inv * q[j]; }
  const float val = neg_dot_i8_order(qp, ix->d.int8_dataset + (size_t)dp * D, D);
  if (ix->d.distance == SO_DOT_PRODUCT) return val;
  const float qn = (float)squared_l2_norm_f64(q, D);
  const float s = qn + ix->d.dp_norms[dp];
  const float t = 2.0f * val;
  return s + t;
}

/* bfloat16 reordering (utils/reordering_helper.cc:745-757, Bfloat16ReorderingHelper::ComputeDistancesForReordering
 * -> DenseDotProductDistanceOneToManyBf16Float / OneToManyBf16FloatSquaredL2): f32 query x bf16 row, f32 FMA. */
static float exact_distance(const so_index* ix, const float* q, uint32_t dp) {
  const uint32_t D = ix->d.d;
  if (!ix->d.dataset && !ix->d.bf16_dataset) return exact_distance_i8(ix, q, dp);
  if (!ix->d.dataset) {
    const int16_t* xb = ix->d.bf16_dataset + (size_t)dp * D;
    return ix->d.distance == SO_DOT_PRODUCT ? neg_dot_bf16_avx2_order(q, xb, D) : sql2_bf16_avx2_order(q, xb, D);
  }
  const float* x = ix->d.dataset + (size_t)dp * D;
  if (ix->d.distance == SO_DOT_PRODUCT) return D < 8 ? neg_dot_small(q, x, D) : neg_dot_avx2_order(q, x, D);
  return D < 8 ? sql2_small(q, x, D) : sql2_avx2_order(q, x, D);
}

int so_exact_distances(const so_index* ix, const float* q, const uint32_t* dps, uint32_t n, float* out) {
  if (!ix->d.dataset && !ix->d.bf16_dataset && !ix->d.int8_dataset) return fail("no dataset");
  for (uint32_t i = 0; i < n; ++i) out[i] = exact_distance(ix, q, dps[i]);
  return 0;
}

typedef struct {
  int k, npre, nover, P;
} sp_t;

static sp_t resolve_params(const so_index* ix, int final_nn, int pre_nn, int leaves) {
  /* scann_ops/cc/scann.cc:406-430 + SetUnspecifiedParametersToDefaults */
  sp_t p;
  const int has_reorder = (ix->d.dataset != NULL || ix->d.bf16_dataset != NULL || ix->d.int8_dataset != NULL) && ix->d.n_blocks != 0;
  p.k = final_nn > 0 ? final_nn : ix->d.default_final_nn;
  if (has_reorder) p.npre = pre_nn > 0 ? pre_nn : ix->d.default_pre_nn;
  else p.npre = p.k;
  /* tree_ah_hybrid_residual.h:263-267 + internal/utils.h:146-157 */
  if (ix->disjoint) p.nover = p.npre;
  else {
    double r = (double)p.npre * (double)ix->d.overretrieve;
    p.nover = r > 2147483647.0 ? 2147483647 : (int)r;
  }
  p.P = leaves > 0 ? leaves : ix->d.default_leaves;
  if ((uint32_t)p.P > ix->d.n_leaves) p.P = (int)ix->d.n_leaves;
  return p;
}

/* one batch, single thread: tree_ah_hybrid_residual.cc:631-786 */
static void search_batch(const so_index* ix, const float* q, uint32_t nq, sp_t sp, int impl,
                         uint64_t** out_keys, size_t* out_n, uint64_t* scan_bytes) {
  const uint32_t L = ix->d.n_leaves, B = ix->d.n_blocks, D = ix->d.d;
  const int P = sp.P;
  int32_t* leaves = (int32_t*)malloc(sizeof(int32_t) * (size_t)nq * P);
  float* biases = (float*)malloc(sizeof(float) * (size_t)nq * P);
  uint8_t* luts = (uint8_t*)aligned_alloc(64, ((size_t)nq * B * 16 + 63) / 64 * 64 + 64);
  float* mults = (float*)malloc(sizeof(float) * nq);
  topn_t* tns = (topn_t*)malloc(sizeof(topn_t) * nq);
  {
    float* ds = (float*)malloc(sizeof(float) * L);
    uint64_t* ks = (uint64_t*)malloc(sizeof(uint64_t) * L);
    float* raw = (float*)malloc(sizeof(float) * B * 16);
    for (uint32_t i = 0; i < nq; ++i) {
      tokenize_one(ix, q + (size_t)i * D, P, ds, ks, leaves + (size_t)i * P, biases + (size_t)i * P);
      lut_one(ix, q + (size_t)i * D, luts + (size_t)i * B * 16, mults + i, raw);
      topn_init(&tns[i], (size_t)sp.nover);
    }
    free(ds); free(ks); free(raw);
  }
  /* InvertCentersToSearch (:610-622) */
  uint32_t* cnt = (uint32_t*)calloc(L + 1, sizeof(uint32_t));
  for (size_t i = 0; i < (size_t)nq * P; ++i) cnt[leaves[i] + 1]++;
  for (uint32_t l = 0; l < L; ++l) cnt[l + 1] += cnt[l];
  uint32_t* lq = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)nq * P);
  float* lb = (float*)malloc(sizeof(float) * (size_t)nq * P);
  uint32_t* cur = (uint32_t*)malloc(sizeof(uint32_t) * L);
  memcpy(cur, cnt, sizeof(uint32_t) * L);
  for (uint32_t i = 0; i < nq; ++i)
    for (int r = 0; r < P; ++r) {
      uint32_t l = (uint32_t)leaves[(size_t)i * P + r];
      lq[cur[l]] = i;
      lb[cur[l]] = ix->d.distance == SO_SQUARED_L2 ? 0.0f : biases[(size_t)i * P + r];
      cur[l]++;
    }
  uint64_t bytes = 0;
  /* The reference visits leaves in descending centre-norm order (:121-143); the exact
   * top-N contract makes the visiting order irrelevant, so plain leaf order is used. */
  for (uint32_t l = 0; l < L; ++l) {
    const uint32_t nql = cnt[l + 1] - cnt[l];
    if (!nql) continue;
    const uint32_t n = ix->leaf_off[l + 1] - ix->leaf_off[l];
    bytes += (uint64_t)nql * ((n + 31) / 32) * 16 * B;
    if (!n) continue;
    if (impl == 0) {
      for (uint32_t e = cnt[l]; e < cnt[l + 1]; ++e)
        scan_leaf_scalar(ix, l, luts + (size_t)lq[e] * B * 16, mults[lq[e]], lb[e], &tns[lq[e]]);
    } else {
      for (uint32_t s = 0; s < nql;) {
        uint32_t left = nql - s;
        int nb = left <= 3 ? (int)left : (left >= 6 ? 3 : (int)(left / 2)); /* :750-755 */
        const uint8_t* la[SO_MAXQ]; float ma[SO_MAXQ], ba[SO_MAXQ]; topn_t* ta[SO_MAXQ];
        for (int j = 0; j < nb; ++j) {
          uint32_t e = cnt[l] + s + (uint32_t)j;
          la[j] = luts + (size_t)lq[e] * B * 16; ma[j] = mults[lq[e]]; ba[j] = lb[e]; ta[j] = &tns[lq[e]];
        }
        scan_leaf_avx2(ix, l, nb, la, ma, ba, ta);
        s += (uint32_t)nb;
      }
    }
  }
  for (uint32_t i = 0; i < nq; ++i) {
    topn_finish(&tns[i]);
    out_keys[i] = tns[i].buf;
    out_n[i] = tns[i].n;
  }
  *scan_bytes += bytes;
  free(cnt); free(lq); free(lb); free(cur); free(leaves); free(biases); free(luts); free(mults); free(tns);
}

/* single_machine_base.cc:569-587 FindNeighborsBatched: search -> reorder -> sort. */
static void finish_query(const so_index* ix, const float* q, sp_t sp, uint64_t* keys, size_t n,
                         uint32_t* out_idx, float* out_dist, int out_k) {
  cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (n + 1));
  uint64_t* k2 = (uint64_t*)malloc(sizeof(uint64_t) * (n + 1));
  for (size_t i = 0; i < n; ++i) {
    c[i].dp = ix->d.distance == SO_SQUARED_L2 ? (uint32_t)keys[i] : ix->leaf_dp[(uint32_t)keys[i]];
    c[i].score = ord2f((uint32_t)(keys[i] >> 32));
  }
  size_t m = n;
  if (!ix->disjoint) m = soar_dedup(c, n, (size_t)sp.npre, k2);
  const int has_reorder = ix->d.dataset != NULL || ix->d.bf16_dataset != NULL || ix->d.int8_dataset != NULL;
  for (size_t i = 0; i < m; ++i) {
    float dist = has_reorder ? exact_distance(ix, q, c[i].dp) : c[i].score;
    k2[i] = ((uint64_t)f2ord(dist) << 32) | c[i].dp;
  }
  /* single_machine_base.cc:872-901 (SortAndDropResults): top-k by (distance, id), sorted */
  size_t kk = (size_t)sp.k < m ? (size_t)sp.k : m;
  select_smallest(k2, m, kk);
  qsort(k2, kk, sizeof(uint64_t), cmp_u64);
  const float mulr = ix->d.distance == SO_DOT_PRODUCT ? -1.0f : 1.0f; /* scann.cc:364-369 */
  for (int i = 0; i < out_k; ++i) {
    if ((size_t)i < kk) {
      out_idx[i] = (uint32_t)k2[i];
      out_dist[i] = mulr * ord2f((uint32_t)(k2[i] >> 32));
    } else { /* scann.h:175-178 */
      out_idx[i] = 0;
      out_dist[i] = NAN;
    }
  }
  free(c); free(k2);
}

int so_search_batched(const so_index* ix, const float* q, uint32_t nq, int final_nn, int pre_nn,
                      int leaves, uint32_t* out_idx, float* out_dist, int out_k, int impl,
                      int threads, int batch) {
  if (!ix->d.n_leaves || !ix->d.n_blocks) return fail("oracle: only tree-AH indexes are searchable here");
  sp_t sp = resolve_params(ix, final_nn, pre_nn, leaves);
  if (sp.k <= 0 || sp.npre <= 0 || sp.P <= 0) return fail("bad search parameters");
  const uint32_t D = ix->d.d;
  uint64_t total_bytes = 0;
  if (threads <= 1) batch = (int)nq;
  if (batch <= 0) batch = 256;
  const uint32_t nbatches = (nq + (uint32_t)batch - 1) / (uint32_t)batch;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1) reduction(+ : total_bytes)
#endif
  for (uint32_t bi = 0; bi < nbatches; ++bi) {
    const uint32_t s = bi * (uint32_t)batch, e = s + (uint32_t)batch < nq ? s + (uint32_t)batch : nq;
    uint64_t** keys = (uint64_t**)malloc(sizeof(uint64_t*) * (e - s));
    size_t* ns = (size_t*)malloc(sizeof(size_t) * (e - s));
    uint64_t bytes = 0;
    search_batch(ix, q + (size_t)s * D, e - s, sp, impl, keys, ns, &bytes);
    for (uint32_t i = s; i < e; ++i) {
      finish_query(ix, q + (size_t)i * D, sp, keys[i - s], ns[i - s], out_idx + (size_t)i * out_k,
                   out_dist + (size_t)i * out_k, out_k);
      free(keys[i - s]);
    }
    free(keys); free(ns);
    total_bytes += bytes;
  }
  g_scan_bytes = total_bytes;
  return 0;
}

int so_candidates(const so_index* ix, const float* q, uint32_t nq, int pre_nn, int leaves, int cap,
                  uint32_t* out_leaf, uint32_t* out_slot, uint32_t* out_dp, float* out_score,
                  int32_t* out_acc, uint32_t* out_count) {
  sp_t sp = resolve_params(ix, -1, pre_nn, leaves);
  uint64_t** keys = (uint64_t**)malloc(sizeof(uint64_t*) * nq);
  size_t* ns = (size_t*)malloc(sizeof(size_t) * nq);
  uint64_t bytes = 0;
  search_batch(ix, q, nq, sp, 0, keys, ns, &bytes);
  g_scan_bytes = bytes;
  uint64_t band = 0;
  const uint32_t L = ix->d.n_leaves, B = ix->d.n_blocks, D = ix->d.d;
  uint8_t* lut = (uint8_t*)malloc((size_t)B * 16);
  float* raw = (float*)malloc(sizeof(float) * B * 16);
  for (uint32_t i = 0; i < nq; ++i) {
    float mult;
    lut_one(ix, q + (size_t)i * D, lut, &mult, raw);
    const float inv = inv_multiplier(ix, mult);
    size_t n = ns[i] < (size_t)cap ? ns[i] : (size_t)cap;
    out_count[i] = (uint32_t)n;
    float last = n ? ord2f((uint32_t)(keys[i][ns[i] - 1] >> 32)) : 0.0f;
    for (size_t j = 0; j < n; ++j) {
      if (ix->d.distance == SO_SQUARED_L2) { /* keys carry the datapoint id, not the slot */
        size_t o = (size_t)i * cap + j;
        out_leaf[o] = 0xFFFFFFFFu; out_slot[o] = 0xFFFFFFFFu; out_acc[o] = 0;
        out_dp[o] = (uint32_t)keys[i][j];
        out_score[o] = ord2f((uint32_t)(keys[i][j] >> 32));
        continue;
      }
      uint32_t gs = (uint32_t)keys[i][j];
      uint32_t lo = 0, hi = L; /* leaf of global slot */
      while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (ix->leaf_off[mid] <= gs) lo = mid; else hi = mid; }
      /* skip empty leaves sharing the same offset */
      while (ix->leaf_off[lo + 1] <= gs) ++lo;
      size_t o = (size_t)i * cap + j;
      out_leaf[o] = lo;
      out_slot[o] = gs - ix->leaf_off[lo];
      out_dp[o] = ix->leaf_dp[gs];
      out_score[o] = ord2f((uint32_t)(keys[i][j] >> 32));
      out_acc[o] = score_slot(lut, ix->slot_codes + (size_t)gs * B, B);
      if (ns[i] == (size_t)sp.nover && out_score[o] > last - inv) band++;
    }
    free(keys[i]);
  }
  g_band = band;
  free(lut); free(raw); free(keys); free(ns);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* bfloat16 brute force (config C3)                                           */
/* ------------------------------------------------------------------------- */


/* brute_force/bfloat16_brute_force.cc:131-147: dist_i = -sum_d q[d] * f32(bf16 x[i][d]) with an f32
 * query and f32 accumulation, over ALL rows, then the k smallest (distance, index).
 * DenseDotProductDistanceOneToManyBf16Float = OneToManyAsymmetricTemplate<dims, 3, .., int16_t> on AVX2
 * (distance_measures/one_to_many/one_to_many_asymmetric_impl.inc:296-353,697-721), the kernel of the int8 path with
 * Bfloat16Decompress on the loads: eight fnmadd lanes over whole groups of 8 dims (HandleXDims<16> is two such steps),
 * one 4-wide step into lanes 0..3, HorizontalSum3X = ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)), the remaining dims one by
 * one with fnmadd on the scalar.  Pinned to the reference's compiled kernel by tests/test_oracle_ref.py.  (The last
 * n mod 3 rows of a call go through ComputeOneToOneScore, :260-289 -- 16-wide groups only, a horizontal sum, every
 * remaining dim on the scalar -- which can differ in the last bit when dims is not a multiple of 16; which rows of a
 * candidate list those are is unspecified, so every row uses the main kernel.) */
static float neg_dot_bf16_avx2_order(const float* q, const int16_t* x, uint32_t n) {
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8)
    for (int l = 0; l < 8; ++l) a[l] = fmaf(-q[j + l], bf16_to_f32(x[j + l]), a[l]);
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) a[l] = fmaf(-q[j + l], bf16_to_f32(x[j + l]), a[l]);
    j += 4;
  }
  float r = ((a[0] + a[4]) + (a[2] + a[6])) + ((a[1] + a[5]) + (a[3] + a[7]));
  for (; j < n; ++j) r = fmaf(-q[j], bf16_to_f32(x[j]), r);
  return r;
}

/* squared-L2 sibling (FusedMultiplyOp<kIsSquaredL2 = true, int16_t>: diff = q - x; acc = fma(diff, diff, acc)) */
static float sql2_bf16_avx2_order(const float* q, const int16_t* x, uint32_t n) {
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint32_t j = 0;
  for (; j + 8 <= n; j += 8)
    for (int l = 0; l < 8; ++l) { const float t = q[j + l] - bf16_to_f32(x[j + l]); a[l] = fmaf(t, t, a[l]); }
  if (j + 4 <= n) {
    for (int l = 0; l < 4; ++l) { const float t = q[j + l] - bf16_to_f32(x[j + l]); a[l] = fmaf(t, t, a[l]); }
    j += 4;
  }
  float r = ((a[0] + a[4]) + (a[2] + a[6])) + ((a[1] + a[5]) + (a[3] + a[7]));
  for (; j < n; ++j) { const float t = q[j] - bf16_to_f32(x[j]); r = fmaf(t, t, r); }
  return r;
}

int so_bruteforce_bf16(const int16_t* db, uint32_t n, uint32_t d, const float* q, uint32_t nq, int k,
                       uint32_t* out_idx, float* out_dist, int threads) {
  if (k <= 0) return fail("k must be positive");
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
#endif
  for (uint32_t i = 0; i < nq; ++i) {
    topn_t tn;
    topn_init(&tn, (size_t)k);
    const float* qi = q + (size_t)i * d;
    for (uint32_t r = 0; r < n; ++r) {
      float dist = neg_dot_bf16_avx2_order(qi, db + (size_t)r * d, d);
      topn_push(&tn, ((uint64_t)f2ord(dist) << 32) | r);
    }
    topn_finish(&tn);
    for (int j = 0; j < k; ++j) {
      if ((size_t)j < tn.n) {
        out_idx[(size_t)i * k + j] = (uint32_t)tn.buf[j];
        out_dist[(size_t)i * k + j] = -ord2f((uint32_t)(tn.buf[j] >> 32));
      } else {
        out_idx[(size_t)i * k + j] = 0;
        out_dist[(size_t)i * k + j] = NAN;
      }
    }
    free(tn.buf);
  }
  return 0;
}

/* Float brute force: BruteForceSearcher<float>::FinishBatchedSearchSimple (brute_force/brute_force.cc:376-393)
 * -> DenseDistanceManyToManyTopK (distance_measures/many_to_many/many_to_many_floating_point.h:119-130,
 * many_to_many_impl.inc:522-567): dist = acc after  acc = 0; acc = fnmadd(q[d], x[d], acc)  sequentially in d
 * (the same kernel as the partitioner's query tokenization), then FastTopNeighbors: the k smallest
 * (distance, index).  Dot product only; API distances are -dist (scann.cc:364-369). */
int so_bruteforce_f32(const float* db, uint32_t n, uint32_t d, const float* q, uint32_t nq, int k, uint32_t* out_idx,
                      float* out_dist, int threads) {
  if (k <= 0) return fail("k must be positive");
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
#endif
  for (uint32_t i = 0; i < nq; ++i) {
    topn_t tn;
    topn_init(&tn, (size_t)k);
    const float* qi = q + (size_t)i * d;
    for (uint32_t r = 0; r < n; ++r) {
      const float* x = db + (size_t)r * d;
      float acc = 0.f;
      for (uint32_t j = 0; j < d; ++j) acc = fmaf(-qi[j], x[j], acc);
      topn_push(&tn, ((uint64_t)f2ord(acc) << 32) | r);
    }
    topn_finish(&tn);
    for (int j = 0; j < k; ++j) {
      if ((size_t)j < tn.n) {
        out_idx[(size_t)i * k + j] = (uint32_t)tn.buf[j];
        out_dist[(size_t)i * k + j] = -ord2f((uint32_t)(tn.buf[j] >> 32));
      } else {
        out_idx[(size_t)i * k + j] = 0;
        out_dist[(size_t)i * k + j] = NAN;
      }
    }
    free(tn.buf);
  }
  return 0;
}

/* Float brute force, squared L2: the same searcher with SquaredL2Distance (brute_force/brute_force.cc:376-393 ->
 * DenseDistanceManyToManyTopK<kIsSquaredL2>): database norms from the transposer, norm = fnmadd(x, x, norm) in
 * dimension order, times -1, and the rows doubled (AugmentWithL2Norms, many_to_many_impl.inc:236-257); query norms
 * from SquaredL2Norm (:417-426, double accumulation); acc = ||x||^2 + ||q||^2, then acc = fnmadd(q[d], 2 x[d], acc)
 * sequentially in d (:527-567).  The distance is acc itself; k smallest (distance, index). */
int so_bruteforce_f32_l2(const float* db, uint32_t n, uint32_t d, const float* q, uint32_t nq, int k, uint32_t* out_idx,
                         float* out_dist, int threads) {
  if (k <= 0) return fail("k must be positive");
  float* xn = (float*)malloc(sizeof(float) * (n ? n : 1));
  for (uint32_t r = 0; r < n; ++r) {
    const float* x = db + (size_t)r * d;
    float a = 0.0f;
    for (uint32_t j = 0; j < d; ++j) a = fmaf(-x[j], x[j], a);
    xn[r] = a * -1.0f;
  }
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
#endif
  for (uint32_t i = 0; i < nq; ++i) {
    topn_t tn;
    topn_init(&tn, (size_t)k);
    const float* qi = q + (size_t)i * d;
    const float qnf = (float)squared_l2_norm_f64(qi, d);
    for (uint32_t r = 0; r < n; ++r) {
      const float* x = db + (size_t)r * d;
      float acc = xn[r] + qnf;
      for (uint32_t j = 0; j < d; ++j) { const float x2 = x[j] * 2.0f; acc = fmaf(-qi[j], x2, acc); }
      topn_push(&tn, ((uint64_t)f2ord(acc) << 32) | r);
    }
    topn_finish(&tn);
    for (int j = 0; j < k; ++j) {
      if ((size_t)j < tn.n) {
        out_idx[(size_t)i * k + j] = (uint32_t)tn.buf[j];
        out_dist[(size_t)i * k + j] = ord2f((uint32_t)(tn.buf[j] >> 32));
      } else {
        out_idx[(size_t)i * k + j] = 0;
        out_dist[(size_t)i * k + j] = NAN;
      }
    }
    free(tn.buf);
  }
  free(xn);
  return 0;
}

/* ========================================================================= */
/* index build, deterministic part: database             */
/* tokenization, SOAR secondary assignment, residuals, AH encoding.            */
/* Given the trained centres and codebook these are pure functions of the      */
/* inputs; the trainers themselves (k-means, codebooks) are random-initialised */
/* and out of scope.                                                            */
/* ========================================================================= */

static __thread uint64_t g_encode_ties;
uint64_t so_last_encode_ties(void) { return g_encode_ties; }

/* KMeansTreePartitioner::TokenizeDatabaseImplFastPath / TokenForDatapointBatchedImpl
 * (partitioning/kmeans_tree_partitioner.cc:561-612,903-923) -> DenseDistanceManyToManyTop1 with
 * SquaredL2Distance (the builder's partitioning distance, scann_builder.py:213-238):
 * acc = ||c||^2 + ||x||^2, then acc = fnmadd(x[dim], 2 c[dim], acc) sequentially in dim
 * (many_to_many_impl.inc:236-257,522-567), ManyToManyTop1Callback keeps the first strict minimum
 * (many_to_many_common.h:176-199) = argmin by (distance, centre index). */
static void l2_center_distances(const float* x, const float* centers_t, const float* cnorm, uint32_t L, uint32_t D, float* out) {
  const float qnf = (float)squared_l2_norm_f64(x, D);  /* SquaredL2Norm of the "query" (a datapoint), see center_distances */
  for (uint32_t l = 0; l < L; ++l) out[l] = cnorm[l] + qnf;
  for (uint32_t k = 0; k < D; ++k) {
    const float nq = -x[k];
    const float* c = centers_t + (size_t)k * L;
    for (uint32_t l = 0; l < L; ++l) { float c2 = c[l] * 2.0f; out[l] = fmaf(nq, c2, out[l]); }
  }
}

static void transpose_centers(const float* centers, uint32_t L, uint32_t D, float* ct, float* cnorm) {
  for (uint32_t l = 0; l < L; ++l)
    for (uint32_t k = 0; k < D; ++k) ct[(size_t)k * L + l] = centers[(size_t)l * D + k];
  for (uint32_t l = 0; l < L; ++l) {
    float a = 0.0f;
    for (uint32_t k = 0; k < D; ++k) { float c = centers[(size_t)l * D + k]; a = fmaf(-c, c, a); }
    cnorm[l] = a * -1.0f;
  }
}

int so_assign_primary(const float* x, uint32_t n, uint32_t d, const float* centers, uint32_t L, int32_t* out_tok,
                      float* out_dist, int threads) {
  if (!L || !d) return fail("so_assign_primary: empty centres");
  float* ct = (float*)malloc(sizeof(float) * (size_t)L * d);
  float* cn = (float*)malloc(sizeof(float) * L);
  transpose_centers(centers, L, d, ct, cn);
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 1 ? threads : 1)
#endif
  {
    float* dist = (float*)malloc(sizeof(float) * L);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (uint32_t i = 0; i < n; ++i) {
      l2_center_distances(x + (size_t)i * d, ct, cn, L, d, dist);
      uint32_t best = 0;
      for (uint32_t l = 1; l < L; ++l) if (dist[l] < dist[best]) best = l;
      out_tok[i] = (int32_t)best;
      if (out_dist) out_dist[i] = dist[best];
    }
    free(dist);
  }
  free(ct); free(cn);
  return 0;
}

/* ComputeNormalizedResidual (partitioning/orthogonality_amplification_utils.h:27-46). */
static void normalized_residual(const float* x, const float* c, uint32_t d, float* out) {
  double sqnorm = 0.0;
  for (uint32_t i = 0; i < d; ++i) {
    out[i] = (float)((double)x[i] - (double)c[i]);
    sqnorm += (double)out[i] * (double)out[i];
  }
  if (sqnorm < 1e-7) { for (uint32_t i = 0; i < d; ++i) out[i] = 0.0f; return; }
  const float inv_norm = (float)(1.0 / sqrt(sqnorm));
  for (uint32_t i = 0; i < d; ++i) out[i] = out[i] * inv_norm;
}

/* KMeansTreePartitioner::OrthogonalityAmplifiedTokenForDatapointBatched
 * (partitioning/kmeans_tree_partitioner.cc:925-997) -> DenseManyToManyOrthogonalityAmplified
 * (distance_measures/many_to_many/many_to_many_impl.inc:729-781), float:
 *   diff = x[dim] - c[dim]; t1 = fma(diff, diff, t1); t2 = fma(diff, rhat[dim], t2)   sequentially in dim
 *   cost = t1 + (lambda * t2) * t2
 * over ALL centres (the primary is not excluded; a datapoint whose secondary equals its primary is
 * simply not spilled, kmeans_tree_partitioner.cc:527-531), first strict minimum. */
int so_assign_soar(const float* x, uint32_t n, uint32_t d, const float* centers, uint32_t L, const int32_t* primary,
                   float lambda, int32_t* out_tok, float* out_cost, int threads) {
  if (!L || !d) return fail("so_assign_soar: empty centres");
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 1 ? threads : 1)
#endif
  {
    float* rhat = (float*)malloc(sizeof(float) * d);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (uint32_t i = 0; i < n; ++i) {
      const float* xi = x + (size_t)i * d;
      normalized_residual(xi, centers + (size_t)primary[i] * d, d, rhat);
      uint32_t best = 0;
      float best_cost = INFINITY;
      for (uint32_t l = 0; l < L; ++l) {
        const float* c = centers + (size_t)l * d;
        float t1 = 0.0f, t2 = 0.0f;
        for (uint32_t k = 0; k < d; ++k) {
          const float diff = xi[k] - c[k];
          t1 = fmaf(diff, diff, t1);
          t2 = fmaf(diff, rhat[k], t2);
        }
        const float lt = lambda * t2;
        const float q = lt * t2;
        const float cost = t1 + q;
        if (cost < best_cost) { best_cost = cost; best = l; }
      }
      out_tok[i] = (int32_t)best;
      if (out_cost) out_cost[i] = best_cost;
    }
    free(rhat);
  }
  return 0;
}

/* DenseSingleAccumulate with l2_distance_internal::Square (utils/reduction.h:357-390,
 * distance_measures/one_to_one/l2_distance.h:57-63,108-111): SquaredL2Norm(DatapointPtr<float>) in double. */
static double squared_l2_norm_f64(const float* v, uint32_t n) {
  double r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  uint32_t i = 0;
  for (; i + 4 <= n; i += 4) {
    r0 += (double)v[i] * (double)v[i];
    r1 += (double)v[i + 1] * (double)v[i + 1];
    r2 += (double)v[i + 2] * (double)v[i + 2];
    r3 += (double)v[i + 3] * (double)v[i + 3];
  }
  r2 += r3;
  if (i + 2 <= n) {
    r0 += (double)v[i] * (double)v[i];
    r1 += (double)v[i + 1] * (double)v[i + 1];
    i += 2;
  }
  r1 += r2;
  if (i < n) r0 += (double)v[i] * (double)v[i];
  return r0 + r1;
}

/* One (datapoint, token) pair.
 *   threshold NaN : Indexer::Hash -> AhImpl::IndexDatapoint (hashes/internal/asymmetric_hashing_impl.cc:199-244):
 *                   per block DenseDistanceOneToMany(SquaredL2) to the 16 centres, std::min_element (first minimum).
 *   otherwise     : Indexer::HashWithNoiseShaping -> AhImpl::IndexDatapointNoiseShaped (:434-503) with
 *                   ComputeResidualStats (:292-340), ComputeParallelCostMultiplier (:258-265),
 *                   InitializeToMinResidualNorm (:342-355), OptimizeSingleSubspace (:376-412); all in double,
 *                   compiled without FMA (no -mfma outside the SCANN_AVX2 functions).
 * `res` is the vector that is hashed (the residual for TreeAHHybridResidual,
 * tree_ah_hybrid_residual.cc:414-428), `orig` the original datapoint.
 * The blocks are visited in descending order of their initial residual norm (ZipSortBranchOptimized with
 * std::greater, :466-476); that sort is not stable, so equal norms (never seen on continuous data; counted in
 * *ties) are ordered here by ascending block index. */
static void encode_one(const float* res, const float* orig, uint32_t D, const float* codebook, uint32_t B, uint32_t S,
                       const int32_t* block_dims, const uint32_t* block_off, double threshold, uint8_t* out,
                       double* stat_norm, double* stat_par, uint64_t* ties) {
  if (isnan(threshold)) {
    float dist[16];
    for (uint32_t b = 0; b < B; ++b) {
      one_to_many(SO_SQUARED_L2, res + block_off[b], codebook + (size_t)b * 16 * S, 16, S, (uint32_t)block_dims[b], dist);
      uint32_t best = 0;
      for (uint32_t c = 1; c < 16; ++c) if (dist[c] < dist[best]) best = c;
      out[b] = (uint8_t)best;
    }
    return;
  }
  /* ComputeResidualStats: chunked norm of the original, then per (block, centre) residual norm and the
   * component of the quantisation residual parallel to the original. */
  double chunked_norm = 0.0;
  for (uint32_t k = 0; k < D; ++k) { const double v = (double)orig[k]; chunked_norm += v * v; }
  chunked_norm = sqrt(chunked_norm);
  const double inv_norm = 1.0 / chunked_norm;
  for (uint32_t b = 0; b < B; ++b) {
    const uint32_t nd = (uint32_t)block_dims[b];
    for (uint32_t c = 0; c < 16; ++c) {
      const float* cen = codebook + ((size_t)b * 16 + c) * S;
      double rn = 0.0, par = 0.0;
      for (uint32_t k = 0; k < nd; ++k) {
        const double rc = (double)res[block_off[b] + k] - (double)cen[k];
        const double sq = rc * rc;
        rn += sq;
        const double p0 = rc * (double)orig[block_off[b] + k];
        const double p1 = p0 * inv_norm;
        par += p1;
      }
      stat_norm[b * 16 + c] = rn;
      stat_par[b * 16 + c] = par;
    }
  }
  /* ComputeParallelCostMultiplier(threshold, SquaredL2Norm(original), dims) */
  const double sqn = squared_l2_norm_f64(orig, D);
  const double t2 = threshold * threshold;
  const double parallel_cost = t2 / sqn;
  const double perpendicular_cost = (1.0 - t2 / sqn) / ((double)D - 1.0);
  const double mult = parallel_cost / perpendicular_cost;
  uint8_t code[256];
  uint16_t order[256];
  double norm0[256];
  for (uint32_t b = 0; b < B; ++b) {
    uint32_t best = 0;
    for (uint32_t c = 1; c < 16; ++c) if (stat_norm[b * 16 + c] < stat_norm[b * 16 + best]) best = c;
    code[b] = (uint8_t)best;
  }
  double par = 0.0;
  for (uint32_t b = 0; b < B; ++b) par += stat_par[b * 16 + code[b]];
  for (uint32_t b = 0; b < B; ++b) { norm0[b] = stat_norm[b * 16 + code[b]]; order[b] = (uint16_t)b; }
  /* descending norm, ties by ascending block (insertion sort: B <= 256) */
  int tie = 0;
  for (uint32_t i = 1; i < B; ++i) {
    const uint16_t v = order[i];
    uint32_t j = i;
    while (j > 0 && norm0[order[j - 1]] < norm0[v]) { order[j] = order[j - 1]; --j; }
    order[j] = v;
  }
  for (uint32_t i = 1; i < B; ++i) if (norm0[order[i]] == norm0[order[i - 1]]) tie = 1;
  if (tie && ties) ++*ties;
  int changes = 1;
  for (int round = 0; changes && round < 10; ++round) {
    changes = 0;
    for (uint32_t i = 0; i < B; ++i) {
      const uint32_t b = order[i];
      const uint8_t cur = code[b];
      const double old_norm = stat_norm[b * 16 + cur], old_par = stat_par[b * 16 + cur];
      uint8_t best = cur;
      double best_delta = 0.0, best_par = par;
      for (uint32_t c = 0; c < 16; ++c) {
        if (c == cur) continue;
        const double d0 = par - old_par;
        const double new_par = d0 + stat_par[b * 16 + c];
        const double a2 = new_par * new_par, b2 = par * par;
        const double par_delta = a2 - b2;
        if (par_delta > 0.0) continue;
        const double norm_delta = stat_norm[b * 16 + c] - old_norm;
        const double perp_delta = norm_delta - par_delta;
        const double m0 = mult * par_delta;
        const double cost_delta = m0 + perp_delta;
        if (cost_delta < best_delta) { best = (uint8_t)c; best_delta = cost_delta; best_par = new_par; }
      }
      if (best != cur) { par = best_par; code[b] = best; changes = 1; }
    }
  }
  for (uint32_t b = 0; b < B; ++b) out[b] = code[b];
}

/* Codes of n (datapoint, token) pairs: row i hashes x[i] - centers[token[i]] (centers != NULL: residual
 * quantisation, TreeAHHybridResidual::ComputeResiduals tree_ah_hybrid_residual.cc:189-212, float subtraction)
 * or x[i] itself (centers == NULL).  out is [n][B], one code per byte. */
int so_encode(const float* x, uint32_t n, uint32_t d, const float* centers, const int32_t* token, const float* codebook,
              uint32_t B, uint32_t S, const int32_t* block_dims_in, double threshold, uint8_t* out, int threads) {
  if (!B || B > 256) return fail("so_encode: 1 <= B <= 256");
  int32_t* bd = (int32_t*)malloc(sizeof(int32_t) * B);
  uint32_t* bo = (uint32_t*)malloc(sizeof(uint32_t) * (B + 1));
  bo[0] = 0;
  for (uint32_t b = 0; b < B; ++b) { bd[b] = block_dims_in ? block_dims_in[b] : (int32_t)S; bo[b + 1] = bo[b] + (uint32_t)bd[b]; }
  if (bo[B] != d) {
    const uint32_t got = bo[B];
    free(bd); free(bo);
    return fail("so_encode: block dims sum to %u, expected %u", got, d);
  }
  uint64_t ties_total = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads > 1 ? threads : 1) reduction(+ : ties_total)
#endif
  {
    float* res = (float*)malloc(sizeof(float) * d);
    double* sn = (double*)malloc(sizeof(double) * B * 16);
    double* sp = (double*)malloc(sizeof(double) * B * 16);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (uint32_t i = 0; i < n; ++i) {
      const float* xi = x + (size_t)i * d;
      const float* r = xi;
      if (centers && token[i] >= 0) {
        const float* c = centers + (size_t)token[i] * d;
        for (uint32_t k = 0; k < d; ++k) res[k] = xi[k] - c[k];
        r = res;
      }
      if (centers && token[i] < 0) { memset(out + (size_t)i * B, 0, B); continue; }
      encode_one(r, xi, d, codebook, B, S, bd, bo, threshold, out + (size_t)i * B, sn, sp, &ties_total);
    }
    free(res); free(sn); free(sp);
  }
  g_encode_ties = ties_total;
  free(bd); free(bo);
  return 0;
}


/* ------------------------------------------------------------------------- */
/* k-means training: Lloyd iterations                    */
/* ------------------------------------------------------------------------- */

/* GmmUtils main loop (utils/gmm_utils.cc:846-915) with UNBALANCED_FLOAT32 assignment and
 * RecomputeCentroidsSimple (:1052-1132): assignment = first minimum of the float many-to-many squared-L2 chain
 * (so_assign_primary); centroid = double sums over the members in index order inside kParallelAggregate = 4 contiguous
 * slices of the training set (one slice when n < 8 k, :1064-1065), slice sums added in slice order into a zeroed
 * accumulator, times double(1.0 / count) (NormalizeCentroid :1038-1048), stored as float for the next assignment.
 * An empty cluster keeps its centre (the reference re-initialises it randomly; see csrc/train.cu).
 * centers: [k][d] in / out; assign_out [n] = the final partition (may be NULL). */
int so_kmeans(const float* x, uint32_t n, uint32_t d, float* centers, uint32_t k, int iterations, int32_t* assign_out,
              uint32_t* empty_out, int threads) {
  if (!n || !d || !k || k > n) return fail("so_kmeans: bad sizes");
  int32_t* assign = (int32_t*)malloc(sizeof(int32_t) * n);
  const uint32_t slices = ((uint64_t)n >= (uint64_t)k * 8) ? 4u : 1u;
  const uint32_t per = (n + slices - 1) / slices;
  double* part = (double*)malloc(sizeof(double) * (size_t)slices * k * d);
  uint32_t* cnt = (uint32_t*)malloc(sizeof(uint32_t) * k);
  uint32_t empty = 0;
  for (int it = 0; it < iterations; ++it) {
    so_assign_primary(x, n, d, centers, k, assign, NULL, threads);
    memset(part, 0, sizeof(double) * (size_t)slices * k * d);
    memset(cnt, 0, sizeof(uint32_t) * k);
    for (uint32_t t = 0; t < slices; ++t) {
      const uint32_t lo = t * per, hi = (lo + per < n) ? lo + per : n;
      for (uint32_t i = lo; i < hi; ++i) {
        double* acc = part + ((size_t)t * k + (uint32_t)assign[i]) * d;
        const float* row = x + (size_t)i * d;
        for (uint32_t j = 0; j < d; ++j) acc[j] += (double)row[j];
        cnt[assign[i]]++;
      }
    }
    empty = 0;
    for (uint32_t c = 0; c < k; ++c) {
      if (!cnt[c]) { ++empty; continue; }
      const double mult = 1.0 / (double)cnt[c];
      for (uint32_t j = 0; j < d; ++j) {
        double sum = 0.0;
        for (uint32_t t = 0; t < slices; ++t) sum += part[((size_t)t * k + c) * d + j];
        centers[(size_t)c * d + j] = (float)(sum * mult);
      }
    }
  }
  if (assign_out) so_assign_primary(x, n, d, centers, k, assign_out, NULL, threads);
  if (empty_out) *empty_out = empty;
  free(assign); free(part); free(cnt);
  return 0;
}
