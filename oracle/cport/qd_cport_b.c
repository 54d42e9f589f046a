/*
 * qd_cport_b.c -- plain-C restatement of the reference's tunnel-coupled ground state ("Path B", what QADAPT's env.step
 * runs): CPU baseline of that path.  TEST INFRASTRUCTURE (oracle/__init__.py).
 *
 * Literal, in the reference's own formulation -- nothing the CUDA kernel does to be fast is used here:
 *   src/qarray_latched/DotArrays/ground_state.py:24-166        orchestration, optional linear capacitance model
 *   src/qarray_latched/DotArrays/charge_states.py:36-88        relaxation: closed form or 50 projected-gradient steps
 *   src/qarray_latched/DotArrays/charge_states.py:135-222      ALL 4^N candidates floor + {-1,0,1,2}^N, each a full
 *                                                              quadratic form; the 32 lowest by (energy, index)
 *   src/qarray_latched/DotArrays/hamiltonian_build.py:12-137   free energies of the kept states, nearest-neighbour hopping
 *   src/qarray_latched/DotArrays/barrier_voltage_model.py      t = tc_base exp(-alpha (vb + Cbg vg))
 *   jnp.linalg.eigh -> here: cyclic Jacobi on the 32 x 32 matrix; <n> = sum_m psi_m^2 n_m
 * Pinned like oracle/path_b.py: tests/test_cport.py checks it against the NumPy oracle and against the fixtures made by
 * the reference itself (tests/golden/ref_*tunnel*.npz).  Parallelism: OpenMP over pixels.
 *
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o libqd_cport_b.so qd_cport_b.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 8
#define MAXV 16
#define M 32

/* cyclic Jacobi eigen-decomposition of the symmetric h[M][M]; eigenvalues in w, eigenvectors in the columns of v */
static void jacobi(double h[M][M], double w[M], double v[M][M]) {
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) v[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < M; ++i) {
      diag += h[i][i] * h[i][i];
      for (int j = i + 1; j < M; ++j) off += h[i][j] * h[i][j];
    }
    if (off <= 1e-30 * (diag + off) || off == 0.0) break;
    for (int p = 0; p < M - 1; ++p)
      for (int q = p + 1; q < M; ++q) {
        const double apq = h[p][q];
        if (apq == 0.0) continue;
        const double theta = (h[q][q] - h[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < M; ++k) {
          const double hkp = h[k][p], hkq = h[k][q];
          h[k][p] = c * hkp - s * hkq;
          h[k][q] = s * hkp + c * hkq;
        }
        for (int k = 0; k < M; ++k) {
          const double hpk = h[p][k], hqk = h[q][k];
          h[p][k] = c * hpk - s * hqk;
          h[q][k] = s * hpk + c * hqk;
        }
        for (int k = 0; k < M; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < M; ++i) w[i] = h[i][i];
}

static void one_pixel(int n, int nv, int g_cnt, const double* cinv, const double* a, const double* cbg, double tc_base,
                      const double* alpha, double vc_alpha, double vc_beta, const double* v, double* n_out,
                      double* gap_out) {
  const int nb = n - 1;
  double g[MAXN], s_c = 1.0, s_g = 1.0;
  if (vc_alpha != 0.0 || vc_beta != 0.0) {
    double m = 0.0;
    for (int k = 0; k < nv; ++k) m += fabs(v[k]);
    m /= (double)nv;
    s_c = 1.0 + vc_alpha * m;
    s_g = 1.0 + vc_beta * m;
  }
  int any_neg = 0;
  for (int i = 0; i < n; ++i) {
    double acc = 0.0;
    for (int k = 0; k < nv; ++k) acc += a[i * nv + k] * v[k];
    g[i] = acc * s_g;
    any_neg |= g[i] < 0.0;
  }
  /* continuous relaxation */
  double nc[MAXN];
  for (int i = 0; i < n; ++i) nc[i] = g[i];
  if (any_neg) {
    double cg[MAXN], x[MAXN], y[MAXN];
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += cinv[i * n + k] * g[k];
      cg[i] = acc / s_c;
      x[i] = g[i] > 0.0 ? g[i] : 0.0;
    }
    for (int it = 0; it < 50; ++it) {
      for (int i = 0; i < n; ++i) {
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += cinv[i * n + k] * x[k];
        const double t = x[i] - 0.1 * (acc / s_c - cg[i]);
        y[i] = t > 0.0 ? t : 0.0;
      }
      memcpy(x, y, sizeof(double) * n);
    }
    memcpy(nc, x, sizeof(double) * n);
  }
  double f[MAXN];
  for (int i = 0; i < n; ++i) f[i] = floor(nc[i] > 0.0 ? nc[i] : 0.0);
  /* every candidate, the reference's index order (last dot fastest); the 32 lowest by (energy, index) */
  const long total = 1L << (2 * n);
  double best_e[M];
  int best_s[M][MAXN];
  for (int m = 0; m < M; ++m) {
    best_e[m] = INFINITY;
    memset(best_s[m], 0, sizeof(best_s[m]));
  }
  for (long idx = 0; idx < total; ++idx) {
    double z[MAXN];
    int cfg[MAXN], valid = 1;
    for (int i = 0; i < n; ++i) {
      const int dg = (int)((idx >> (2 * (n - 1 - i))) & 3);
      cfg[i] = (int)f[i] + dg - 1;
      valid &= cfg[i] >= 0;
      z[i] = (double)cfg[i] - g[i];
    }
    if (!valid) continue;
    double e = 0.0;
    for (int i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int k = 0; k < n; ++k) acc += cinv[i * n + k] * z[k];
      e += z[i] * acc;
    }
    if (!(e < best_e[M - 1])) continue;            /* strict: equal energies keep the earlier index */
    int pos = M - 1;
    while (pos > 0 && e < best_e[pos - 1]) {
      best_e[pos] = best_e[pos - 1];
      memcpy(best_s[pos], best_s[pos - 1], sizeof(best_s[pos]));
      --pos;
    }
    best_e[pos] = e;
    memcpy(best_s[pos], cfg, sizeof(int) * n);
  }
  /* tunnel couplings */
  double t[MAXN];
  for (int d = 0; d < nb; ++d) {
    double tt = tc_base;
    if (cbg != NULL && nv > g_cnt) {
      double vb = v[g_cnt + d];
      for (int k = 0; k < g_cnt; ++k) vb += cbg[d * g_cnt + k] * v[k];
      tt *= exp(-alpha[d] * vb);
    }
    t[d] = tt;
  }
  /* Hamiltonian */
  double h[M][M], w[M], vec[M][M];
  for (int i = 0; i < M; ++i) {
    for (int j = 0; j < M; ++j) h[i][j] = 0.0;
    double z[MAXN], e = 0.0;
    for (int k = 0; k < n; ++k) z[k] = (double)best_s[i][k] - g[k];
    for (int k = 0; k < n; ++k) {
      double acc = 0.0;
      for (int l = 0; l < n; ++l) acc += cinv[k * n + l] * z[l];
      e += z[k] * acc;
    }
    h[i][i] = e / s_c;
  }
  for (int i = 0; i < M; ++i)
    for (int j = 0; j < M; ++j) {
      if (i == j) continue;
      for (int d = 0; d < nb; ++d) {
        int ok_f = 1, ok_b = 1;
        for (int k = 0; k < n; ++k) {
          const int diff = best_s[j][k] - best_s[i][k];
          const int ef = (k == d) ? -1 : (k == d + 1) ? 1 : 0;
          ok_f &= diff == ef;
          ok_b &= diff == -ef;
        }
        const double nf = (double)best_s[i][d], nt = (double)best_s[i][d + 1];
        if (ok_f) h[i][j] += -t[d] * sqrt(nf * (nt + 1.0));
        if (ok_b) h[i][j] += -t[d] * sqrt(nt * (nf + 1.0));
      }
    }
  jacobi(h, w, vec);
  int i0 = 0;
  for (int i = 1; i < M; ++i) if (w[i] < w[i0]) i0 = i;
  double w1 = INFINITY;
  for (int i = 0; i < M; ++i) if (i != i0 && w[i] < w1) w1 = w[i];
  for (int k = 0; k < n; ++k) {
    double acc = 0.0;
    for (int m = 0; m < M; ++m) acc += vec[m][i0] * vec[m][i0] * (double)best_s[m][k];
    n_out[k] = acc;
  }
  if (gap_out) *gap_out = w1 - w[i0];
}

/* n_pts voltage points v[n_pts][nv] of ONE device -> <n>[n_pts][n] (and the spectral gap per point) */
int qd_cport_tunnel_points(long n_pts, int n, int nv, int g_cnt, const double* cinv, const double* a, const double* cbg,
                           double tc_base, const double* alpha, double vc_alpha, double vc_beta, const double* v,
                           double* n_out, double* gap_out, int threads) {
  if (n < 2 || n > MAXN || nv > MAXV) return -1;
  (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads > 0 ? threads : 1)
#endif
  for (long p = 0; p < n_pts; ++p)
    one_pixel(n, nv, g_cnt, cinv, a, cbg, tc_base, alpha, vc_alpha, vc_beta, v + p * nv, n_out + p * n,
              gap_out ? gap_out + p : NULL);
  return 0;
}
