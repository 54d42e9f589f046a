/*
 * qd_cport.c -- plain-C restatement of the reference's Path A scan (CPU baseline; TEST INFRASTRUCTURE, see
 * oracle/__init__.py).  PARITY UNPINNED, same restatement decisions as oracle/path_a.py, oracle/latching.py,
 * oracle/noise.py, oracle/sensor.py, and validated against them in tests/test_oracle.py.
 *
 * It deliberately keeps the REFERENCE FORMULATION: every candidate's free energy is the full quadratic form
 * (n - cgd v)^T cdd_inv (n - cgd v) (src/qarray_latched/functions.py:30-34), the sensor evaluates the 11 full-system
 * free energies (src/qarray_latched/DotArrays/TunnelCoupledChargeSensed.py:356-378), latching and telegraph noise are
 * sequential loops along the scan.  Parallelism = OpenMP over scans / rows, like the upstream Rust core's rayon loop
 * over voltage points.
 *
 * Build: gcc -O3 -march=x86-64-v3 -fopenmp -fPIC -shared -o libqd_cport.so qd_cport.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/qdsim.h"

#define MAXN QD_MAX_DOTS

/* ---- Philox4x32-10, contract of oracle/philox.py ---- */
static void philox(uint64_t seed, uint64_t index, uint32_t purpose, uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = purpose, c3 = 0;
  for (int r = 0; r < 10; ++r) {
    if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u24(uint32_t w) { return (double)(w >> 8) * (1.0 / 16777216.0); }

/* ---- exact relaxation: monotone active set, dense Gaussian elimination on the clamped block ---- */
static void relax(int n, const double* g, const double* cdd, double* nc) {
  int act[MAXN], any = 0;
  for (int i = 0; i < n; ++i) { act[i] = g[i] < 0; any |= act[i]; nc[i] = g[i]; }
  if (!any) return;
  double w[MAXN];
  for (int round = 0; round <= n; ++round) {
    int idx[MAXN], k = 0;
    for (int i = 0; i < n; ++i) if (act[i]) idx[k++] = i;
    double m[MAXN][MAXN + 1];
    for (int a = 0; a < k; ++a) {
      for (int b = 0; b < k; ++b) m[a][b] = cdd[idx[a] * n + idx[b]];
      m[a][k] = -g[idx[a]];
    }
    for (int p = 0; p < k; ++p) {
      for (int a = p + 1; a < k; ++a) {
        double f = m[a][p] / m[p][p];
        for (int b = p; b <= k; ++b) m[a][b] -= f * m[p][b];
      }
    }
    double mu[MAXN];
    for (int a = k - 1; a >= 0; --a) {
      double s = m[a][k];
      for (int b = a + 1; b < k; ++b) s -= m[a][b] * mu[b];
      mu[a] = s / m[a][a];
    }
    int changed = 0;
    for (int i = 0; i < n; ++i) {
      double s = g[i];
      for (int a = 0; a < k; ++a) s += cdd[i * n + idx[a]] * mu[a];
      if (act[i]) s = 0.0;
      w[i] = s;
      if (s < 0 && !act[i]) { act[i] = 1; changed = 1; }
    }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) nc[i] = w[i] > 0 ? w[i] : 0.0;
}

static double energy(int n, const double* conf, const double* g, const double* cinv) {
  double e = 0.0;
  for (int i = 0; i < n; ++i) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s += cinv[i * n + j] * (conf[j] - g[j]);
    e += (conf[i] - g[i]) * s;
  }
  return e;
}

/* ground state of one pixel; returns occupations in nd */
static void ground_state(int n, int alg, const double* g, const double* cinv, const double* cdd, double thr, int maxc,
                         double kT, double* nd, double* margin) {
  /* *margin <- second-lowest minus lowest candidate energy (oracle/path_a.py::_select): the tie detector of the
   * parity checks (an exact tie has no defined winner across two summation orders) */
  double conf[MAXN], best_conf[MAXN], best = INFINITY, second = INFINITY;
  if (alg == QD_ALG_BRUTE_FORCE) {
    long total = 1;
    for (int i = 0; i < n; ++i) total *= (maxc + 1);
    double z = 0.0, acc[MAXN] = {0};
    for (int pass = 0; pass < (kT > 0 ? 2 : 1); ++pass) {
      for (long c = 0; c < total; ++c) {
        long t = c;
        for (int i = n - 1; i >= 0; --i) { conf[i] = (double)(t % (maxc + 1)); t /= (maxc + 1); }
        double e = energy(n, conf, g, cinv);
        if (pass == 0) { if (e < best) { second = best; best = e; memcpy(best_conf, conf, sizeof(double) * n); } else if (e < second) second = e; }
        else { double wgt = exp(-(e - best) / kT); z += wgt; for (int i = 0; i < n; ++i) acc[i] += wgt * conf[i]; }
      }
    }
    for (int i = 0; i < n; ++i) nd[i] = kT > 0 ? acc[i] / z : best_conf[i];
    if (margin) *margin = second - best;
    return;
  }
  double nc[MAXN], fl[MAXN];
  int fixed[MAXN], fixv[MAXN];
  relax(n, g, cdd, nc);
  for (int i = 0; i < n; ++i) {
    fl[i] = floor(nc[i]);
    fixed[i] = 0; fixv[i] = 0;
    if (alg == QD_ALG_THRESHOLDED) {
      double frac = nc[i] - fl[i];
      if (!(fabs(frac - 0.5) < thr / 2.0)) { fixed[i] = 1; fixv[i] = (int)(floor(nc[i] + 0.5) - fl[i]); }
    }
  }
  double z = 0.0, acc[MAXN] = {0};
  for (int pass = 0; pass < (kT > 0 ? 2 : 1); ++pass) {
    for (unsigned c = 0; c < (1u << n); ++c) {
      int ok = 1;
      for (int i = 0; i < n; ++i) {
        int d = (c >> (n - 1 - i)) & 1;
        if (fixed[i] && d != fixv[i]) ok = 0;
        conf[i] = fl[i] + d;
      }
      if (!ok) continue;
      double e = energy(n, conf, g, cinv);
      if (pass == 0) { if (e < best) { second = best; best = e; memcpy(best_conf, conf, sizeof(double) * n); } else if (e < second) second = e; }
      else { double wgt = exp(-(e - best) / kT); z += wgt; for (int i = 0; i < n; ++i) acc[i] += wgt * conf[i]; }
    }
  }
  for (int i = 0; i < n; ++i) nd[i] = kT > 0 ? acc[i] / z : best_conf[i];
  if (margin) *margin = second - best;
}

static void scan_rows(const qd_scan* s, int n, int nv, int alg, unsigned flags, const double* cinv, const double* cdd,
                      const double* cinv_full, const double* cgd_full, const qd_env_params* p, int row0, int row1,
                      float* z_out, double* n_out, double* margin_out) {
  const int d = n + 1, nx = s->nx;
  const double kT = (flags & QD_FLAG_THERMAL) ? p->kT : 0.0;
  const int latch = (flags & QD_FLAG_LATCH) && p->latching;
  const int noise = flags & QD_FLAG_NOISE;
  const int rad_mode = (flags & QD_FLAG_RADIAL) ? s->rad_mode : 0;
  const int carry = flags & QD_FLAG_CARRY_ROWS;
  double held[MAXN], heldk[MAXN];
  int have_held = 0, tele = 0, tele_init = 0;
  const double tot = p->tele_p01 + p->tele_p10;
  for (int iy = row0; iy < row1; ++iy) {
    if (!carry) { have_held = 0; tele_init = 0; }
    for (int ix = 0; ix < nx; ++ix) {
      const uint64_t pix = (uint64_t)iy * nx + ix;
      uint32_t w[4];
      philox(s->seed, pix, 0, w);
      const double u1 = ((double)(w[0] >> 8) + 1.0) / 16777216.0, u2 = u24(w[1]);
      const double rr = sqrt(-2.0 * log(u1)), ang = 2.0 * M_PI * u2;
      const double z_white = rr * cos(ang), z_rad = rr * sin(ang), u_latch = u24(w[2]), u_tele = u24(w[3]);
      double v[QD_MAX_VOLT], g[MAXN], nd[MAXN];
      for (int k = 0; k < nv; ++k) v[k] = (s->v0[k] + ix * s->dx[k]) + iy * s->dy[k];
      for (int i = 0; i < n; ++i) { double a = 0; for (int k = 0; k < nv; ++k) a += cgd_full[i * nv + k] * v[k]; g[i] = a; }
      double mg = INFINITY;
      ground_state(n, alg, g, cinv, cdd, p->threshold, p->max_charge_carriers, kT, nd, margin_out ? &mg : NULL);
      if (margin_out) margin_out[s->pix_offset + (int64_t)pix] = mg;
      if (latch) {
        double key[MAXN];
        for (int i = 0; i < n; ++i) key[i] = (flags & QD_FLAG_LATCH_EXACT) ? nd[i] : floor(nd[i] + 0.5);
        if (!have_held) { memcpy(held, nd, sizeof(double) * n); memcpy(heldk, key, sizeof(double) * n); have_held = 1; }
        else {
          int nd_ = 0, d0 = -1, d1 = -1;
          for (int i = 0; i < n; ++i) if (key[i] != heldk[i]) { if (nd_ == 0) d0 = i; else d1 = i; ++nd_; }
          int accept = 1;
          if (nd_ == 1) accept = u_latch < p->p_leads[d0];
          else if (nd_ == 2) accept = u_latch < p->p_inter[d0 * QD_MAX_DOTS + d1];
          if (accept) { memcpy(held, nd, sizeof(double) * n); memcpy(heldk, key, sizeof(double) * n); }
          else memcpy(nd, held, sizeof(double) * n);
        }
      }
      double noise_in = 0.0, noise_out = 0.0;
      if (noise) {
        if (flags & QD_FLAG_WHITE_ON_OUTPUT) noise_out = p->white_amp * z_white; else noise_in = p->white_amp * z_white;
        if (p->tele_amp != 0.0) {
          if (!tele_init) {
            if (carry) tele = 0;
            else { uint32_t wr[4]; philox(s->seed, (uint64_t)iy, 1, wr); tele = (tot > 0 && u24(wr[0]) < p->tele_p01 / tot) ? 1 : 0; }
            tele_init = 1;
          }
          if (u_tele < (tele ? p->tele_p10 : p->tele_p01)) tele ^= 1;
          noise_in += p->tele_amp * tele;
        }
      }
      /* sensor: eleven full-system free energies, first differences, ten Lorentzians */
      double vdash[MAXN + 1], nf[MAXN + 1], f[11];
      for (int i = 0; i < d; ++i) { double a = 0; for (int k = 0; k < nv; ++k) a += cgd_full[i * nv + k] * v[k]; vdash[i] = a; }
      const double ns = nearbyint(vdash[n]);
      for (int i = 0; i < n; ++i) nf[i] = nd[i];
      for (int k = -5; k <= 5; ++k) {
        nf[n] = ns + k + noise_in;
        double e = 0.0;
        for (int i = 0; i < d; ++i) { double sacc = 0; for (int j = 0; j < d; ++j) sacc += cinv_full[i * d + j] * (nf[j] - vdash[j]); e += (nf[i] - vdash[i]) * sacc; }
        f[k + 5] = e;
      }
      double z = 0.0;
      for (int k = 0; k < 10; ++k) { double x = (f[k + 1] - f[k]) / s->peak_width; z += 1.0 / (x * x + 1.0); }
      z += noise_out;
      if (rad_mode == 1) {
        double vx = s->rad_x0 + ix * s->rad_dx, vy = s->rad_y0 + iy * s->rad_dy;
        double amp = s->rad_alpha * (sqrt(vx * vx + vy * vy) - s->rad_zero_radius);
        amp = amp < 0 ? 0 : (amp > s->rad_max_amp ? s->rad_max_amp : amp);
        z += z_rad * amp;
      } else if (rad_mode == 2) z = z_rad;
      const int64_t o = s->pix_offset + (int64_t)pix;
      if (z_out) z_out[o] = (float)z;
      if (n_out) for (int i = 0; i < n; ++i) n_out[o * n + i] = nd[i];
    }
  }
}

/* Simulate n_scan scans on `threads` OpenMP threads.  Arrays are per env, same layout as qd_set_models. */
int qd_cport_scans(int n_scan, const qd_scan* scans, int n_dot, int n_volt, int algorithm, unsigned flags,
                   const double* cdd_inv_gs, const double* cdd_gs, const double* cdd_inv_full, const double* cgd_full,
                   const qd_env_params* params, float* z_out, double* n_out, double* margin_out, int threads) {
  const int n = n_dot, d = n_dot + 1, nv = n_volt;
  if (n < 1 || n > MAXN || nv > QD_MAX_VOLT) return -1;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  const int carry = flags & QD_FLAG_CARRY_ROWS;
  if (carry || n_scan >= 4 * (threads > 0 ? threads : 1)) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n_scan; ++i) {
      const qd_scan* s = scans + i;
      const int e = s->env_id;
      scan_rows(s, n, nv, algorithm, flags, cdd_inv_gs + (size_t)e * n * n, cdd_gs + (size_t)e * n * n,
                cdd_inv_full + (size_t)e * d * d, cgd_full + (size_t)e * d * nv, params + e, 0, s->ny, z_out, n_out, margin_out);
    }
  } else {
    for (int i = 0; i < n_scan; ++i) {
      const qd_scan* s = scans + i;
      const int e = s->env_id;
#pragma omp parallel for schedule(dynamic, 1)
      for (int iy = 0; iy < s->ny; ++iy)
        scan_rows(s, n, nv, algorithm, flags, cdd_inv_gs + (size_t)e * n * n, cdd_gs + (size_t)e * n * n,
                  cdd_inv_full + (size_t)e * d * d, cgd_full + (size_t)e * d * nv, params + e, iy, iy + 1, z_out, n_out, margin_out);
    }
  }
  return 0;
}
