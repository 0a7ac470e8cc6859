/* CPU ORACLE (test infrastructure, NOT product code) -- C restatement of the ALS
 * half-step the reference reaches through pyspark at src/als_model.py:62.
 *
 * The arithmetic lives in Spark MLlib 3.5.1 (ml/recommendation/ALS.scala), which
 * is not under /root/reference and cannot run in this image: PARITY UNPINNED
 * (see oracle/als_oracle.py header).  This file restates, per destination row:
 *
 *   NormalEquation.add      : copyToDouble(y); dspr('U', k, c, y, ata); daxpy(k, b, y, atb)
 *   computeYtY              : dspr over every source row (implicit mode)
 *   CholeskySolver.solve    : ata[diag] += lambda*n ; dppsv('U', k, 1, ata, atb) ; (float) x
 *
 * packed-upper storage: ata[i + j*(j+1)/2], i <= j.  All accumulation in double.
 * Threads: OpenMP over destination rows (Spark's own parallelism is over
 * destination blocks; rows are independent so the result does not depend on it).
 *
 * Built by oracle/Makefile into oracle/_build/libals_oracle.so.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* dspr, uplo='U', incx=1: ap += alpha * x x^T on the packed upper triangle */
static void dspr_upper(int k, double alpha, const double *x, double *ap) {
  int kk = 0;
  for (int j = 0; j < k; ++j) {
    const double t = alpha * x[j];
    double *col = ap + kk;
    for (int i = 0; i <= j; ++i) col[i] += x[i] * t;
    kk += j + 1;
  }
}

/* dpptrf('U') + dpptrs: A = U^T U on packed upper, then solve U^T U x = b in place.
 * Returns 0, or j+1 if the leading minor of order j+1 is not positive definite. */
static int dppsv_upper(int k, double *ap, double *b) {
  /* factorisation, column by column (left-looking, as LAPACK dpptrf 'U') */
  int jj = 0;
  for (int j = 0; j < k; ++j) {
    double *cj = ap + jj; /* column j: cj[0..j] */
    /* solve U(0:j,0:j)^T * v = A(0:j, j) (dtpsv 'U','T','N') */
    int ii = 0;
    for (int i = 0; i < j; ++i) {
      const double *ci = ap + ii;
      double s = cj[i];
      for (int p = 0; p < i; ++p) s -= ci[p] * cj[p];
      cj[i] = s / ci[i];
      ii += i + 1;
    }
    double d = cj[j];
    for (int p = 0; p < j; ++p) d -= cj[p] * cj[p];
    if (!(d > 0.0)) return j + 1;
    cj[j] = sqrt(d);
    jj += j + 1;
  }
  /* U^T y = b */
  jj = 0;
  for (int j = 0; j < k; ++j) {
    const double *cj = ap + jj;
    double s = b[j];
    for (int p = 0; p < j; ++p) s -= cj[p] * b[p];
    b[j] = s / cj[j];
    jj += j + 1;
  }
  /* U x = y */
  for (int j = k - 1; j >= 0; --j) {
    const double *cj = ap + (size_t)j * (j + 1) / 2;
    b[j] /= cj[j];
    const double xj = b[j];
    for (int p = 0; p < j; ++p) b[p] -= cj[p] * xj;
  }
  return 0;
}

/* Y^T Y in packed upper, double (computeYtY). */
void oracle_gram_packed(const float *src, int64_t n_src, int k, double *ap_out) {
  const int np = k * (k + 1) / 2;
  memset(ap_out, 0, sizeof(double) * np);
#pragma omp parallel
  {
    double *loc = (double *)calloc(np, sizeof(double));
    double *y = (double *)malloc(sizeof(double) * k);
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n_src; ++i) {
      for (int f = 0; f < k; ++f) y[f] = (double)src[i * k + f];
      dspr_upper(k, 1.0, y, loc);
    }
#pragma omp critical
    for (int p = 0; p < np; ++p) ap_out[p] += loc[p];
    free(loc);
    free(y);
  }
}

/* One ALS half-step over rows [row_begin,row_end) of a CSR matrix.
 * Returns 0 on success, otherwise 1 + index of the first row whose system was not SPD. */
int64_t oracle_als_half_step(const int64_t *rowptr, const int32_t *colidx, const float *vals,
                             int64_t row_begin, int64_t row_end, const float *src, int64_t n_src,
                             int k, double reg, int implicit, double alpha, float *dst) {
  const int np = k * (k + 1) / 2;
  double *yty = NULL;
  if (implicit) {
    yty = (double *)malloc(sizeof(double) * np);
    oracle_gram_packed(src, n_src, k, yty);
  }
  int64_t bad = 0;
#pragma omp parallel
  {
    double *ata = (double *)malloc(sizeof(double) * np);
    double *atb = (double *)malloc(sizeof(double) * k);
    double *y = (double *)malloc(sizeof(double) * k);
#pragma omp for schedule(dynamic, 16)
    for (int64_t j = row_begin; j < row_end; ++j) {
      const int64_t lo = rowptr[j], hi = rowptr[j + 1];
      float *x = dst + j * k;
      if (hi == lo) {
        for (int f = 0; f < k; ++f) x[f] = 0.0f;
        continue;
      }
      if (implicit) memcpy(ata, yty, sizeof(double) * np);
      else memset(ata, 0, sizeof(double) * np);
      memset(atb, 0, sizeof(double) * k);
      int64_t n = 0;
      for (int64_t p = lo; p < hi; ++p) {
        const float *ys = src + (int64_t)colidx[p] * k;
        for (int f = 0; f < k; ++f) y[f] = (double)ys[f];
        const double r = (double)vals[p];
        if (implicit) {
          const double c1 = alpha * fabs(r);
          dspr_upper(k, c1, y, ata);
          if (r > 0.0) {
            const double bw = 1.0 + c1;
            for (int f = 0; f < k; ++f) atb[f] += bw * y[f];
            ++n;
          }
        } else {
          dspr_upper(k, 1.0, y, ata);
          if (r != 0.0)
            for (int f = 0; f < k; ++f) atb[f] += r * y[f];
          ++n;
        }
      }
      const double lam = reg * (double)n;
      int d = 0;
      for (int f = 0; f < k; ++f) {
        d += f; /* index of (f,f): f + f(f+1)/2 */
        ata[d + f] += lam;
      }
      const int info = dppsv_upper(k, ata, atb);
      if (info != 0) {
#pragma omp critical
        if (bad == 0 || j + 1 < bad) bad = j + 1;
        for (int f = 0; f < k; ++f) x[f] = 0.0f;
      } else {
        for (int f = 0; f < k; ++f) x[f] = (float)atb[f];
      }
    }
    free(ata);
    free(atb);
    free(y);
  }
  free(yty);
  return bad;
}

/* ALSModel.transform dot: sequential float accumulation. */
void oracle_als_predict(const float *X, const float *Y, int k, const int32_t *users,
                        const int32_t *items, int64_t n, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; ++p) {
    const float *a = X + (int64_t)users[p] * k;
    const float *b = Y + (int64_t)items[p] * k;
    float s = 0.0f;
    for (int f = 0; f < k; ++f) s += a[f] * b[f];
    out[p] = s;
  }
}

/* Dense scoring for the hybrid CPU baseline: per user, both score rows, min-max
 * blend (hybrid_system.py:66-72) and a partial top-k with (score desc, item asc)
 * ordering.  fp32 dots accumulated in double, blend in double like the numpy
 * reference path. */
void oracle_hybrid_topk(const float *Ua, const float *Ia, int ka, const float *Ut, const float *It,
                        int kt, int64_t n_users, int64_t n_items, double w_als, double w_tt,
                        int topk, int32_t *out_idx, double *out_score) {
#pragma omp parallel
  {
    double *sa = (double *)malloc(sizeof(double) * n_items);
    double *st = (double *)malloc(sizeof(double) * n_items);
#pragma omp for schedule(dynamic, 1)
    for (int64_t u = 0; u < n_users; ++u) {
      double mna = INFINITY, mxa = -INFINITY, mnt = INFINITY, mxt = -INFINITY;
      for (int64_t i = 0; i < n_items; ++i) {
        float a = 0.0f, t = 0.0f;
        const float *ua = Ua + u * ka, *ia = Ia + i * ka;
        for (int f = 0; f < ka; ++f) a += ua[f] * ia[f];
        const float *ut = Ut + u * kt, *it = It + i * kt;
        for (int f = 0; f < kt; ++f) t += ut[f] * it[f];
        sa[i] = a; st[i] = t;
        if (a < mna) mna = a; if (a > mxa) mxa = a;
        if (t < mnt) mnt = t; if (t > mxt) mxt = t;
      }
      const double sca = (mxa > mna) ? 1.0 / (mxa - mna) : 1.0;
      const double sct = (mxt > mnt) ? 1.0 / (mxt - mnt) : 1.0;
      int32_t *oi = out_idx + u * topk;
      double *os = out_score + u * topk;
      int cnt = 0;
      for (int64_t i = 0; i < n_items; ++i) {
        const double s = w_als * ((sa[i] - mna) * sca) + w_tt * ((st[i] - mnt) * sct);
        if (cnt == topk && !(s > os[cnt - 1])) continue;
        int pos = cnt < topk ? cnt : topk - 1;
        while (pos > 0 && s > os[pos - 1]) {
          os[pos] = os[pos - 1]; oi[pos] = oi[pos - 1]; --pos;
        }
        os[pos] = s; oi[pos] = (int32_t)i;
        if (cnt < topk) ++cnt;
      }
      for (int p = cnt; p < topk; ++p) { oi[p] = -1; os[p] = -INFINITY; }
    }
    free(sa);
    free(st);
  }
}

/* bench.py --impl reference under torch.distributed.run inherits OMP_NUM_THREADS=1: the reference arm asks for every
 * host core explicitly. */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
