/* mpm_oracle.c -- C99 + OpenMP restatement of the taichi_elements MLS-MPM
 * substep (CPU).  TEST INFRASTRUCTURE ONLY: used by tests/ to cross-check the
 * NumPy oracle and at sizes NumPy cannot reach, and by bench.py as the timed
 * CPU baseline ("port": the reference's ti.cpu path cannot run here, Taichi is
 * not installable).  Never linked or loaded by the shipped package.
 *
 * PARITY UNPINNED (see oracle/mpm_oracle.py): no Taichi, no golden vectors in
 * the reference.  This file follows /root/reference/engine/mpm_solver.py
 * line by line (line numbers cited inline); the SVD is an independent
 * one-sided Jacobi (Hestenes) so that it cross-checks the CUDA kernel's
 * two-sided Jacobi rather than sharing its code.
 *
 * Like Taichi's CPU back end (no block-local storage), P2G scatters with
 * atomics straight into the global grid; the grid is dense over the box the
 * caller passes (cells lo..hi, global signed indices), which for the bounded
 * scenes is the whole res^d domain.
 *
 * Arrays are SoA-free "AoS per field": x[n][d], v[n][d], F[n][d*d] row-major,
 * C[n][d*d], Jp[n], material[n].
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int dim;
  int res[3];
  int grid_size, padding, quant, support_plasticity;   /* quant: bit-packed x / v / F storage (3D), see q_* below */
  float dx, inv_dx, p_vol, p_mass, mu_0, lambda_0, alpha, sand_coef, water_density, inv_dx2, four_inv_dx;
  float gravity[3];
} oracle_params;

typedef struct {
  int kind, surface; /* 0 bbox (a[0]=unbounded), 1 sphere, 2 plane */
  float a[3], b[3], r2, friction;
} oracle_collider;

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* explicit thread count: launchers such as torchrun export OMP_NUM_THREADS=1 */
void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- ti.svd restated ----------------------------------------------------- */
static void svd2(const float* F, float* U, float* sig, float* V) {
  /* Taichi's closed form, SURVEY.md Appendix B */
  float a = F[0] + F[3], b = F[2] - F[1];
  float s = 1.0f / sqrtf(a * a + b * b);
  float rc = a * s, rs = b * s;
  float S00 = rc * F[0] + rs * F[2], S01 = rc * F[1] + rs * F[3], S11 = -rs * F[1] + rc * F[3];
  float c = 1.0f, sn = 0.0f, s1 = S00, s2 = S11;
  if (fabsf(S01) >= 1e-5f) {
    float tau = 0.5f * (S00 - S11);
    float w = sqrtf(tau * tau + S01 * S01);
    float t = tau > 0.0f ? S01 / (tau + w) : S01 / (tau - w);
    c = 1.0f / sqrtf(t * t + 1.0f);
    sn = -t * c;
    s1 = c * c * S00 - 2.0f * c * sn * S01 + sn * sn * S11;
    s2 = sn * sn * S00 + 2.0f * c * sn * S01 + c * c * S11;
  }
  float v00, v01, v10, v11;
  if (s1 < s2) { sig[0] = s2; sig[1] = s1; v00 = -sn; v01 = c; v10 = -c; v11 = -sn; }
  else { sig[0] = s1; sig[1] = s2; v00 = c; v01 = sn; v10 = -sn; v11 = c; }
  V[0] = v00; V[1] = v01; V[2] = v10; V[3] = v11;
  U[0] = rc * v00 - rs * v10; U[1] = rc * v01 - rs * v11;
  U[2] = rs * v00 + rc * v10; U[3] = rs * v01 + rc * v11;
}

static float det3(const float* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}

static void svd3(const float* F, float* U, float* sig, float* V) {
  /* one-sided Jacobi: rotate column pairs of B (=F V) until orthogonal */
  float B[9];
  memcpy(B, F, sizeof B);
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0f : 0.0f;
  for (int sweep = 0; sweep < 8; ++sweep) {
    float off = 0.0f;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        float al = 0, be = 0, ga = 0;
        for (int k = 0; k < 3; ++k) { al += B[k * 3 + p] * B[k * 3 + p]; be += B[k * 3 + q] * B[k * 3 + q]; ga += B[k * 3 + p] * B[k * 3 + q]; }
        off += fabsf(ga);
        if (fabsf(ga) <= 1e-12f * sqrtf(al * be) || ga == 0.0f) continue;
        float zeta = (be - al) / (2.0f * ga);
        float t = (zeta >= 0 ? 1.0f : -1.0f) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
        float c = 1.0f / sqrtf(1.0f + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          float bp = B[k * 3 + p], bq = B[k * 3 + q];
          B[k * 3 + p] = c * bp - s * bq; B[k * 3 + q] = s * bp + c * bq;
          float vp = V[k * 3 + p], vq = V[k * 3 + q];
          V[k * 3 + p] = c * vp - s * vq; V[k * 3 + q] = s * vp + c * vq;
        }
      }
    if (off == 0.0f) break;
  }
  float nrm[3];
  for (int j = 0; j < 3; ++j) nrm[j] = sqrtf(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
  /* sort columns by descending norm */
  int idx[3] = {0, 1, 2};
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2 - i; ++j)
      if (nrm[idx[j]] < nrm[idx[j + 1]]) { int t = idx[j]; idx[j] = idx[j + 1]; idx[j + 1] = t; }
  float Vs[9], Us[9];
  for (int j = 0; j < 3; ++j) {
    int c = idx[j];
    sig[j] = nrm[c];
    for (int k = 0; k < 3; ++k) { Vs[k * 3 + j] = V[k * 3 + c]; Us[k * 3 + j] = nrm[c] > 0 ? B[k * 3 + c] / nrm[c] : 0.0f; }
  }
  /* complete U when rank deficient: last column = cross of the first two */
  if (!(sig[2] > 1e-20f * sig[0])) {
    Us[2] = Us[3] * Us[7] - Us[6] * Us[4];
    Us[5] = Us[6] * Us[1] - Us[0] * Us[7];
    Us[8] = Us[0] * Us[4] - Us[3] * Us[1];
  }
  /* convention: U, V rotations; sign of det F on the last singular value */
  if (det3(Us) < 0) { Us[2] = -Us[2]; Us[5] = -Us[5]; Us[8] = -Us[8]; sig[2] = -sig[2]; }
  if (det3(Vs) < 0) { Vs[2] = -Vs[2]; Vs[5] = -Vs[5]; Vs[8] = -Vs[8]; sig[2] = -sig[2]; }
  memcpy(U, Us, sizeof Us);
  memcpy(V, Vs, sizeof Vs);
}

void oracle_svd(int dim, const float* F, int n, float* U, float* sig, float* V) {
  for (int i = 0; i < n; ++i) {
    if (dim == 2) svd2(F + 4 * i, U + 4 * i, sig + 2 * i, V + 4 * i);
    else svd3(F + 9 * i, U + 9 * i, sig + 3 * i, V + 9 * i);
  }
}

#define MM(D, A, B, Cm) for (int i_ = 0; i_ < D; ++i_) for (int j_ = 0; j_ < D; ++j_) { float s_ = 0; for (int k_ = 0; k_ < D; ++k_) s_ += A[i_ * D + k_] * B[k_ * D + j_]; Cm[i_ * D + j_] = s_; }
#define MMT(D, A, B, Cm) for (int i_ = 0; i_ < D; ++i_) for (int j_ = 0; j_ < D; ++j_) { float s_ = 0; for (int k_ = 0; k_ < D; ++k_) s_ += A[i_ * D + k_] * B[j_ * D + k_]; Cm[i_ * D + j_] = s_; }
#define USVT(D, U, s, V, Cm) for (int i_ = 0; i_ < D; ++i_) for (int j_ = 0; j_ < D; ++j_) { float a_ = 0; for (int k_ = 0; k_ < D; ++k_) a_ += U[i_ * D + k_] * s[k_] * V[j_ * D + k_]; Cm[i_ * D + j_] = a_; }

/* per-particle part of p2g, engine/mpm_solver.py:506-574 */
static void particle_update(const oracle_params* P, int D, float dt, int mat, float* F, const float* C, float* Jp,
                            float* affine, float* mass_out) {
  const int DD = D * D;
  float Fin[9], A[9], Fn[9], U[9], V[9], sig[3], stress[9];
  if (mat == 0) {                                              /* :508-511 */
    for (int i = 0; i < DD; ++i) Fin[i] = 0;
    for (int i = 0; i < D; ++i) Fin[i * D + i] = 1;
    if (P->support_plasticity) Fin[0] = *Jp;
  } else memcpy(Fin, F, DD * sizeof(float));
  for (int i = 0; i < DD; ++i) A[i] = dt * C[i];
  for (int i = 0; i < D; ++i) A[i * D + i] += 1.0f;
  MM(D, A, Fin, Fn);                                           /* :513 */
  float h = 1.0f;                                              /* :515-521 */
  if (P->support_plasticity && mat != 0) h = expf(10.0f * (1.0f - *Jp));
  if (mat == 1) h = 0.3f;
  float mu = P->mu_0 * h, la = P->lambda_0 * h;
  if (mat == 0) mu = 0.0f;
  if (D == 2) svd2(Fn, U, sig, V); else svd3(Fn, U, sig, V);   /* :525 */
  float J = 1.0f;
  if (mat != 3) {                                              /* :527-536 */
    for (int d = 0; d < D; ++d) {
      float ns = sig[d];
      if (mat == 2) ns = fminf(fmaxf(sig[d], 1.0f - 2.5e-2f), 1.0f + 4.5e-3f);
      if (P->support_plasticity) *Jp *= sig[d] / ns;
      sig[d] = ns;
      J *= ns;
    }
  }
  if (mat == 0) {                                              /* :537-542 */
    for (int i = 0; i < DD; ++i) Fn[i] = 0;
    for (int i = 0; i < D; ++i) Fn[i * D + i] = 1;
    Fn[0] = J;
    if (P->support_plasticity) *Jp = J;
  } else if (mat == 2) {                                       /* :543-545 */
    USVT(D, U, sig, V, Fn);
  }
  for (int i = 0; i < DD; ++i) stress[i] = 0;
  if (mat != 3) {                                              /* :549-551 */
    float R[9], T[9];
    MMT(D, U, V, R);
    for (int i = 0; i < DD; ++i) T[i] = 2.0f * mu * (Fn[i] - R[i]);
    MMT(D, T, Fn, stress);
    for (int i = 0; i < D; ++i) stress[i * D + i] += la * J * (J - 1.0f);
  } else if (P->support_plasticity) {                          /* :553-566, sand_projection :321-342 */
    float eps[3], eh[3], tr = 0, nrm = 0;
    for (int i = 0; i < D; ++i) { eps[i] = logf(fmaxf(fabsf(sig[i]), 1e-4f)); tr += eps[i]; }
    tr += *Jp;
    for (int i = 0; i < D; ++i) { eh[i] = eps[i] - tr / (float)D; nrm += eh[i] * eh[i]; }
    nrm = sqrtf(nrm) + 1e-20f;
    if (tr >= 0.0f) { *Jp = tr; for (int i = 0; i < D; ++i) sig[i] = 1.0f; }
    else {
      *Jp = 0.0f;
      float dg = nrm + P->sand_coef * tr * P->alpha;
      for (int i = 0; i < D; ++i) sig[i] = expf(eps[i] - fmaxf(0.0f, dg) / nrm * eh[i]);
    }
    USVT(D, U, sig, V, Fn);
    float ls[3], center[3], lsum = 0, T[9];
    for (int i = 0; i < D; ++i) { ls[i] = logf(sig[i]); lsum += ls[i]; }
    for (int i = 0; i < D; ++i) center[i] = 2.0f * P->mu_0 * ls[i] * (1.0f / sig[i]) + P->lambda_0 * lsum * (1.0f / sig[i]);
    USVT(D, U, center, V, T);
    MMT(D, T, Fn, stress);
  }
  memcpy(F, Fn, DD * sizeof(float));                           /* :567 */
  float scale = -dt * P->p_vol * 4.0f * P->inv_dx2;            /* :569 */
  float mass = P->p_mass;
  if (mat == 0) mass *= P->water_density;                      /* :571-573 */
  for (int i = 0; i < DD; ++i) affine[i] = scale * stress[i] + mass * C[i];
  *mass_out = mass;
}

void oracle_particle_update(const oracle_params* P, float dt, int n, const int* mat, float* F, const float* C,
                            float* Jp, float* affine, float* mass) {
  const int DD = P->dim * P->dim;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) particle_update(P, P->dim, dt, mat[i], F + (size_t)DD * i, C + (size_t)DD * i, Jp + i, affine + (size_t)DD * i, mass + i);
}

/* Quantised particle storage (quant=True in 3D, engine/mpm_solver.py:106-114, 216-247): a store into x (3 x fixed 21 bit,
 * max 2.0), v (3 x 19-bit fractions with one shared 7-bit exponent) or F (9 x fixed 16 bit, max F_bound + 0.1) rounds to the
 * field's grid, and later loads see the rounded value.  Taichi's quantised types are not in the reference tree [EXT]; this
 * restates their definitions (as oracle/quant_oracle.py does, which the CUDA codecs are tested against bit for bit). */
static float q_fixed(float x, float max_value, int bits) {
  const float inv = (float)(1 << (bits - 1)) / max_value, scale = max_value / (float)(1 << (bits - 1));
  const float lim = (float)((1 << (bits - 1)) - 1);
  float t = x * inv;
  if (!(t > -lim)) t = -lim;            /* also NaN */
  if (t > lim) t = lim;
  const float r = truncf(t + (t < 0.0f ? -0.5f : 0.5f));   /* round half away from zero */
  return fminf(fmaxf(r, -lim), lim) * scale;
}
static void q_shared_exp3(float* v) {
  const float a = fmaxf(fabsf(v[0]), fmaxf(fabsf(v[1]), fabsf(v[2])));
  int e = -64;
  if (a > 0.0f) {
    int ex;
    frexpf(a, &ex);                     /* a = f 2^ex, f in [0.5, 1): floor(log2 a) = ex - 1 */
    e = ex - 1;
    if (e < -64) e = -64;
    if (e > 63) e = 63;
  }
  const float inv = ldexpf(1.0f, 17 - e), s = ldexpf(1.0f, e - 17), lim = (float)((1 << 18) - 1);
  for (int d = 0; d < 3; ++d) {
    float t = v[d] * inv;
    t = fminf(fmaxf(t, -lim), lim);
    const float r = truncf(t + (t < 0.0f ? -0.5f : 0.5f));
    v[d] = fminf(fmaxf(r, -lim), lim) * s;
  }
}

static inline void atomic_addf(float* p, float v) {
#pragma omp atomic
  *p += v;
}

/* One substep (:789-797) on a dense grid box [lo,hi) of global signed cell
 * indices; grid_mv[cells][dim], grid_m[cells] are caller scratch.  Returns 0,
 * or -1 if a particle's stencil leaves the box. */
int oracle_substep(const oracle_params* P, float dt, int64_t n, float* x, float* v, float* F, float* C, float* Jp,
                   const int* mat, const int* lo, const int* hi, float* gv, float* gm, const oracle_collider* cols,
                   int ncol) {
  const int D = P->dim, DD = D * D;
  int ext[3] = {1, 1, 1};
  size_t ncell = 1;
  for (int d = 0; d < D; ++d) { ext[d] = hi[d] - lo[d]; ncell *= (size_t)ext[d]; }
  int bad = 0;
  /* grid.deactivate_all() :789 */
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < ncell; ++i) { gm[i] = 0; for (int d = 0; d < D; ++d) gv[i * D + d] = 0; }
  /* p2g :487-584 */
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (int64_t p = 0; p < n; ++p) {
    int base[3]; float fx[3], w[3][3];
    for (int d = 0; d < D; ++d) {
      float xs = x[p * D + d] * P->inv_dx;
      base[d] = (int)floorf(xs - 0.5f);
      fx[d] = xs - (float)base[d];
      w[0][d] = 0.5f * (1.5f - fx[d]) * (1.5f - fx[d]);
      w[1][d] = 0.75f - (fx[d] - 1.0f) * (fx[d] - 1.0f);
      w[2][d] = 0.5f * (fx[d] - 0.5f) * (fx[d] - 0.5f);
      if (base[d] < lo[d] || base[d] + 3 > hi[d]) bad = 1;
    }
    float aff[9], mass;
    particle_update(P, D, dt, mat[p], F + p * DD, C + p * DD, Jp + p, aff, &mass);
    if (P->quant && D == 3)                                      /* self.F[p] = F rounds (:567); the stress used the unrounded F */
      for (int i = 0; i < DD; ++i) F[p * DD + i] = q_fixed(F[p * DD + i], 4.1f, 16);
    if (bad) continue;
    const int K = D == 3 ? 27 : 9;
    for (int o = 0; o < K; ++o) {
      int off[3] = {D == 3 ? o / 9 : o / 3, D == 3 ? (o / 3) % 3 : o % 3, o % 3};
      float dpos[3], wt = 1.0f;
      size_t cell = 0;
      for (int d = 0; d < D; ++d) {
        dpos[d] = ((float)off[d] - fx[d]) * P->dx;
        wt *= w[off[d]][d];
        cell = cell * (size_t)ext[d] + (size_t)(base[d] + off[d] - lo[d]);
      }
      for (int r = 0; r < D; ++r) {
        float a = 0;
        for (int c = 0; c < D; ++c) a += aff[r * D + c] * dpos[c];
        atomic_addf(&gv[cell * D + r], wt * (mass * v[p * D + r] + a));
      }
      atomic_addf(&gm[cell], wt * mass);
    }
  }
  if (bad) return -1;
  /* grid_normalization_and_gravity :586-598, grid_postprocess :600-687 */
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < ncell; ++i) {
    int I[3] = {0, 0, 0};
    size_t r = i;
    for (int d = D - 1; d >= 0; --d) { I[d] = (int)(r % (size_t)ext[d]) + lo[d]; r /= (size_t)ext[d]; }
    float* vv = gv + i * D;
    float m = gm[i];
    if (m > 0) {
      float inv = 1.0f / m;
      for (int d = 0; d < D; ++d) vv[d] = inv * vv[d] + dt * P->gravity[d];
    }
    for (int c = 0; c < ncol; ++c) {
      const oracle_collider* col = cols + c;
      if (col->kind == 0) {
        for (int d = 0; d < D; ++d) {
          int l = col->a[0] != 0 ? -P->grid_size / 2 + P->padding : P->padding;
          int h = col->a[0] != 0 ? P->grid_size / 2 - P->padding : P->res[d] - P->padding;
          if (I[d] < l && vv[d] < 0) vv[d] = 0;
          if (I[d] >= h && vv[d] > 0) vv[d] = 0;
        }
      } else if (col->kind == 1) {
        float off[3], nsq = 0;
        for (int d = 0; d < D; ++d) { off[d] = (float)I[d] * P->dx - col->a[d]; nsq += off[d] * off[d]; }
        if (nsq < col->r2) {
          if (col->surface == 0) { for (int d = 0; d < D; ++d) vv[d] = 0; }
          else {
            float invn = 1.0f / (sqrtf(nsq) + 1e-5f), nrm[3], nc = 0;
            for (int d = 0; d < D; ++d) { nrm[d] = off[d] * invn; nc += nrm[d] * vv[d]; }
            float k = col->surface == 1 ? nc : fminf(nc, 0.0f);
            for (int d = 0; d < D; ++d) vv[d] -= nrm[d] * k;
          }
        }
      } else {
        float dotn = 0;
        for (int d = 0; d < D; ++d) dotn += ((float)I[d] * P->dx - col->a[d]) * col->b[d];
        if (dotn < 0) {
          if (col->surface == 0) { for (int d = 0; d < D; ++d) vv[d] = 0; }
          else {
            float nc = 0, nsq = 0;
            for (int d = 0; d < D; ++d) nc += vv[d] * col->b[d];
            float k = col->surface == 1 ? nc : fminf(nc, 0.0f);
            for (int d = 0; d < D; ++d) { vv[d] -= col->b[d] * k; nsq += vv[d] * vv[d]; }
            float norm = sqrtf(nsq);
            if (nc < 0 && norm > 1e-30f) {
              float sc = fmaxf(0.0f, norm + nc * col->friction), invn = 1.0f / norm;
              for (int d = 0; d < D; ++d) vv[d] = vv[d] * invn * sc;
            }
          }
        }
      }
    }
  }
  /* g2p :694-724 */
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < n; ++p) {
    int base[3]; float fx[3], w[3][3];
    for (int d = 0; d < D; ++d) {
      float xs = x[p * D + d] * P->inv_dx;
      base[d] = (int)floorf(xs - 0.5f);
      fx[d] = xs - (float)base[d];
      w[0][d] = 0.5f * (1.5f - fx[d]) * (1.5f - fx[d]);
      w[1][d] = 0.75f - (fx[d] - 1.0f) * (fx[d] - 1.0f);
      w[2][d] = 0.5f * (fx[d] - 0.5f) * (fx[d] - 0.5f);
    }
    float nv[3] = {0, 0, 0}, nC[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int K = D == 3 ? 27 : 9;
    for (int o = 0; o < K; ++o) {
      int off[3] = {D == 3 ? o / 9 : o / 3, D == 3 ? (o / 3) % 3 : o % 3, o % 3};
      float dpos[3], wt = 1.0f;
      size_t cell = 0;
      for (int d = 0; d < D; ++d) {
        dpos[d] = (float)off[d] - fx[d];
        wt *= w[off[d]][d];
        cell = cell * (size_t)ext[d] + (size_t)(base[d] + off[d] - lo[d]);
      }
      const float* g = gv + cell * D;
      for (int r = 0; r < D; ++r) {
        nv[r] += wt * g[r];
        for (int c = 0; c < D; ++c) nC[r * D + c] += P->four_inv_dx * wt * (g[r] * dpos[c]);
      }
    }
    if (mat[p] != 4) {                                         /* :722-724 */
      if (P->quant && D == 3) q_shared_exp3(nv);               /* self.v[p] = new_v rounds; the advection reads it back */
      for (int d = 0; d < D; ++d) { v[p * D + d] = nv[d]; x[p * D + d] += dt * nv[d]; }
      if (P->quant && D == 3)
        for (int d = 0; d < D; ++d) x[p * D + d] = q_fixed(x[p * D + d], 2.0f, 21);
      for (int i = 0; i < DD; ++i) C[p * DD + i] = nC[i];
    }
  }
  return 0;
}
