// crm_esmt.cpp -- TEST ORACLE: Explicit Scalar Momentum Transport, the reference's -DSP_ESMT option (scalar_momentum.F90).
// The GCM wind is carried as two extra, non-conserved scalars u_esmt / v_esmt that are advected and diffused like every other scalar
// and feel an estimate of the pressure-gradient force diagnosed from w * d<u_esmt>/dz by a spectral Poisson solve in x.
//   scalar_momentum_tend   scalar_momentum.F90:119-151
//   scalar_momentum_pgf    scalar_momentum.F90:156-320
//   esmt_fft_forward/backward  scalar_momentum.F90:325-423 (FFTPACK 5.1 rfft1f / rfft1b, fftpack5_1d.F90:386-560)
//   advection with the positivity offset   advect_all_scalars.F90:74-88;   diffusion   SGS_TKE/sgs.F90:442-448
// The transforms are restated as the real DFT pair that FFTPACK 5.1 computes (forward scaled by 1/n resp. 2/n, backward a plain
// sum; forward o backward = identity), evaluated directly: R(1) = mean, R(2j) = (2/n) sum r cos, R(2j+1) = (2/n) sum r sin,
// R(n) = (1/n) sum (-1)^x r.  The sign convention of the sine coefficients cancels in the solve (every coefficient is solved with the
// same real matrix and transformed back with the matching basis function).
#include <cmath>
#include <vector>

#include "crm_routines.h"

namespace orc {

// advect_all_scalars.F90:74-88
void esmt_advect(const Geo &g, Crm &c) {
  double mn = c.u_esmt.d[0];
  for (double v : c.u_esmt.d) mn = dmin(mn, v);
  for (double v : c.v_esmt.d) mn = dmin(mn, v);
  const double off = std::fabs(mn) + R4(50.);
  for (double &v : c.u_esmt.d) v = v + off;
  for (double &v : c.v_esmt.d) v = v + off;
  P1 d1, d2;
  d1.alloc(1, g.nz); d2.alloc(1, g.nz);
  advect_scalar(g, c, c.u_esmt, d1, d2);
  advect_scalar(g, c, c.v_esmt, d1, d2);
  for (double &v : c.u_esmt.d) v = v - off;
  for (double &v : c.v_esmt.d) v = v - off;
}

// SGS_TKE/sgs.F90:442-448 (fluxb = fluxt = 0)
void esmt_diffuse(const Geo &g, Crm &c, const Step &s) {
  const std::vector<double> zero((size_t)g.nx * g.ny, 0.);
  P1 d1, d2;
  d1.alloc(1, g.nz); d2.alloc(1, g.nz);
  diffuse_scalar(g, c, s, c.u_esmt, zero, zero, d1, d2);
  diffuse_scalar(g, c, s, c.v_esmt, zero, zero, d1, d2);
}

// scalar_momentum.F90:156-320 for one scalar; tend(i,j,k) returned through `f` updated in place by the caller
static void scalar_momentum_pgf(const Geo &g, const Crm &c, const F3 &u_s, F3 &tend) {
  const int nx = g.nx, ny = g.ny, nzm = g.nzm;
  const double pi = (double)(2.f * asinf(1.0f));   // pi = 2.*asin(1.0): REAL(4) arithmetic stored in a real(crm_rknd) (scalar_momentum.F90:347)
  std::vector<double> dz(nzm + 2), avg(nzm + 1), shr(nzm + 1), a(nzm + 1), b(nzm + 1), cc(nzm + 1), rhs(nzm + 1), k_arr(nx + 1);
  std::vector<double> w_i((size_t)nx * (nzm + 1)), w_hat((size_t)nx * (nzm + 1)), pgf_hat((size_t)(nx + 1) * (nzm + 1)), ct(nx), st(nx);
  auto WI = [&](int i, int k) -> double & { return w_i[(size_t)(k - 1) * nx + (i - 1)]; };
  auto WH = [&](int i, int k) -> double & { return w_hat[(size_t)(k - 1) * nx + (i - 1)]; };
  auto PH = [&](int i, int k) -> double & { return pgf_hat[(size_t)(k - 1) * (nx + 1) + i]; };
  for (int q = 0; q < nx; q++) { ct[q] = std::cos(2. * 3.14159265358979323846 * q / nx); st[q] = std::sin(2. * 3.14159265358979323846 * q / nx); }
  for (int k = 1; k <= nzm; k++) dz[k] = c.zi(k + 1) - c.zi(k);
  dz[nzm + 1] = dz[nzm];
  // wavenumbers (esmt_fft_forward, :367-380)
  const int nh = (nx % 2 == 0) ? nx / 2 - 1 : (nx - 1) / 2;
  k_arr[1] = 0.;
  for (int j = 1; j <= nh; j++) { k_arr[2 * j] = 2. * pi * (double)j / ((double)nx * g.dx); k_arr[2 * j + 1] = k_arr[2 * j]; }
  if (nx % 2 == 0) k_arr[nx] = 2. * pi / (2. * g.dx);
  for (int j = 1; j <= ny; j++) {
    for (int k = 1; k <= nzm; k++) {
      double sm = 0.;
      for (int i = 1; i <= nx; i++) { sm = sm + u_s(i, j, k); WI(i, k) = (c.w(i, j, k) + c.w(i, j, k + 1)) / 2.0; }
      avg[k] = sm / (double)nx;
      shr[k] = 0.;
    }
    shr[1] = (avg[2] - avg[1]) / (c.z(2) - c.z(1));
    for (int k = 2; k <= nzm - 1; k++) shr[k] = (avg[k + 1] - avg[k - 1]) / (c.z(k + 1) - c.z(k - 1));
    // forward transform of every level
    for (int k = 1; k <= nzm; k++) {
      double s0 = 0., sn = 0.;
      for (int i = 1; i <= nx; i++) { s0 += WI(i, k); sn += ((i - 1) & 1) ? -WI(i, k) : WI(i, k); }
      WH(1, k) = s0 / (double)nx;
      for (int m = 1; m <= nh; m++) {
        double sc = 0., ss = 0.;
        for (int i = 1; i <= nx; i++) { const int q = ((i - 1) * m) % nx; sc += WI(i, k) * ct[q]; ss += WI(i, k) * st[q]; }
        WH(2 * m, k) = 2. * sc / (double)nx; WH(2 * m + 1, k) = 2. * ss / (double)nx;
      }
      if (nx % 2 == 0) WH(nx, k) = sn / (double)nx;
    }
    for (auto &v : pgf_hat) v = 0.;
    for (int i = 2; i <= nx; i++) {
      const double k2 = k_arr[i] * k_arr[i];
      for (int k = 1; k <= nzm; k++) {
        a[k] = dz[k + 1] / (dz[k + 1] + dz[k]);
        b[k] = -0.5 * 1.25 * k2 * dz[k] * dz[k + 1] - 1.0;   // 1.25: crude 2-D vs 3-D updraft geometry factor (:232)
        cc[k] = dz[k] / (dz[k + 1] + dz[k]);
        rhs[k] = k2 * WH(i, k) * shr[k] * dz[k] * dz[k + 1];
      }
      b[1] = b[1] + a[1]; a[1] = 0.;
      b[nzm] = b[nzm] + cc[nzm]; cc[nzm] = 0.;
      for (int k = 1; k <= nzm - 1; k++) {
        b[k + 1] = b[k + 1] - a[k + 1] / b[k] * cc[k];
        rhs[k + 1] = rhs[k + 1] - a[k + 1] / b[k] * rhs[k];
      }
      rhs[nzm] = rhs[nzm] / b[nzm];
      for (int k = nzm - 1; k >= 1; k--) rhs[k] = (rhs[k] - cc[k] * rhs[k + 1]) / b[k];
      for (int k = 1; k <= nzm; k++) PH(i, k) = rhs[k];
    }
    if (nx % 2 == 0) for (int k = 1; k <= nzm; k++) PH(nx, k) = PH(nx, k) / 2.0;   // "something about the Nyquist freq" (:283-289)
    // backward transform + tendency
    for (int k = 1; k <= nzm; k++)
      for (int i = 1; i <= nx; i++) {
        double v = PH(1, k);
        for (int m = 1; m <= nh; m++) { const int q = ((i - 1) * m) % nx; v += PH(2 * m, k) * ct[q] + PH(2 * m + 1, k) * st[q]; }
        if (nx % 2 == 0) v += ((i - 1) & 1) ? -PH(nx, k) : PH(nx, k);
        tend(i, j, k) = (k == 1) ? 0. : -1.0 * v * c.rho(k);
      }
  }
}

// scalar_momentum.F90:119-151
void scalar_momentum_tend(const Geo &g, Crm &c, const Step &s) {
  F3 tu, tv;
  tu.alloc(1, g.nx, 1, g.ny, g.nzm); tv.alloc(1, g.nx, 1, g.ny, g.nzm);
  scalar_momentum_pgf(g, c, c.u_esmt, tu);
  scalar_momentum_pgf(g, c, c.v_esmt, tv);
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        c.u_esmt(i, j, k) = c.u_esmt(i, j, k) + tu(i, j, k) * s.dtn;
        c.v_esmt(i, j, k) = c.v_esmt(i, j, k) + tv(i, j, k) * s.dtn;
      }
}

}  // namespace orc

// ---- KAT hook (tests/test_second_restatement.py): scalar_momentum_pgf for one 2-D CRM; arrays (nx, nz[m]) with i fastest ----
extern "C" int crm_oracle_test_esmt_pgf(int nx, int nzm, double dx, const double *zi, const double *z, const double *rho, const double *w,
                                        const double *u_s, double *tend) {
  using namespace orc;
  Geo g;
  g.nx = nx; g.ny = 1; g.nzm = nzm; g.nz = nzm + 1; g.dx = dx; g.dy = dx; g.YES3D = 0; g.RUN3D = false;
  Crm c;
  c.zi.alloc(1, nzm + 1); c.z.alloc(1, nzm + 1); c.rho.alloc(1, nzm);
  for (int k = 1; k <= nzm + 1; k++) c.zi(k) = zi[k - 1];
  for (int k = 1; k <= nzm; k++) { c.z(k) = z[k - 1]; c.rho(k) = rho[k - 1]; }
  c.w.alloc(1, nx, 1, 1, nzm + 1);
  F3 us, td;
  us.alloc(1, nx, 1, 1, nzm); td.alloc(1, nx, 1, 1, nzm);
  for (int k = 1; k <= nzm + 1; k++) for (int i = 1; i <= nx; i++) c.w(i, 1, k) = w[(size_t)(k - 1) * nx + (i - 1)];
  for (int k = 1; k <= nzm; k++) for (int i = 1; i <= nx; i++) us(i, 1, k) = u_s[(size_t)(k - 1) * nx + (i - 1)];
  scalar_momentum_pgf(g, c, us, td);
  for (int k = 1; k <= nzm; k++) for (int i = 1; i <= nx; i++) tend[(size_t)(k - 1) * nx + (i - 1)] = td(i, 1, k);
  return 0;
}
