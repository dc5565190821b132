// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference CRM (SAM 6.10.4 as embedded in
// E3SM-Project/ACME-ECP, components/cam/src/physics/crm).  Nothing in the product path (acme-ecp_b200/)
// may include, link or call this.  PARITY UNPINNED: the reference ships no golden vectors for crm() and
// no Fortran compiler exists in this environment, so this restatement is pinned only by the analytic /
// invariant known-answer tests in tests/test_oracle_*.py.
//
// Shared helpers: field containers with Fortran-style index ranges, REAL(4)-literal emulation, constants.
#pragma once
#include <vector>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include "../include/crm_b200.h"

// Reference sources are compiled WITHOUT -r8: an unsuffixed literal such as 0.61 is REAL(4) and is
// converted to double when it meets a double operand.  R4(0.61) reproduces that.
#define R4(x) ((double)(x##f))

namespace orc {

// ---- constants: cime/src/share/util/shr_const_mod.F90:22-54 via CRM/params.F90:49-68 -------------
constexpr double SHR_G      = 9.80616;
constexpr double SHR_BOLTZ  = 1.38065e-23;
constexpr double SHR_AVOGAD = 6.02214e26;
constexpr double SHR_RGAS   = SHR_AVOGAD * SHR_BOLTZ;
constexpr double SHR_MWDAIR = 28.966;
constexpr double SHR_MWWV   = 18.016;
constexpr double cp    = 1.00464e3;
constexpr double ggr   = SHR_G;
constexpr double lcond = 2.501e6;
constexpr double lfus  = 3.337e5;
constexpr double lsub  = lcond + lfus;
constexpr double rgas  = SHR_RGAS / SHR_MWDAIR;
constexpr double rv    = SHR_RGAS / SHR_MWWV;
const double diffelq = R4(2.21e-05);   // params.F90:58-60 (REAL(4) literals)
const double therco  = R4(2.40e-02);
const double muelq   = R4(1.717e-05);
constexpr double fac_cond = lcond / cp;
constexpr double fac_fus  = lfus / cp;
constexpr double fac_sub  = lsub / cp;
const double pi_crm = R4(3.141592653589793);   // params.F90:69: literal without kind -> REAL(4)
const double epsv   = R4(0.61);                // setparm.F90:87

// ---- micro_params.F90:12-65 (all REAL(4) literals promoted) --------------------------------------
const double rhor = R4(1000.), rhos = R4(100.), rhog = R4(400.);
const double tbgmin = R4(253.16), tbgmax = R4(273.16), tprmin = R4(268.16), tprmax = R4(283.16);
const double tgrmin = R4(223.16), tgrmax = R4(283.16);
const double a_rain = R4(842.), b_rain = R4(0.8), a_snow = R4(4.84), b_snow = R4(0.25);
const double a_grau = R4(94.5), b_grau = R4(0.5);
const double qcw0 = R4(1.e-3), qci0 = R4(1.e-4), alphaelq = R4(1.e-3), betaelq = R4(1.e-3);
const double erccoef = R4(1.0), esccoef = R4(1.0), esicoef = R4(0.1), egccoef = R4(1.0), egicoef = R4(0.1);
const double nzeror = R4(8.e6), nzeros = R4(3.e6), nzerog = R4(4.e6);
const double qp_threshold = R4(1.e-8);

// 3-D field with Fortran index ranges  f(ilo:ihi, jlo:jhi, 1:nk)
struct F3 {
  int ilo = 1, ihi = 0, jlo = 1, jhi = 0, nk = 0, sx = 0, sxy = 0;
  std::vector<double> d;
  void alloc(int ilo_, int ihi_, int jlo_, int jhi_, int nk_, double v = 0.0) {
    ilo = ilo_; ihi = ihi_; jlo = jlo_; jhi = jhi_; nk = nk_;
    sx = ihi - ilo + 1; sxy = sx * (jhi - jlo + 1);
    d.assign((size_t)sxy * nk, v);
  }
  inline double &operator()(int i, int j, int k) { return d[(size_t)(k - 1) * sxy + (j - jlo) * sx + (i - ilo)]; }
  inline double operator()(int i, int j, int k) const { return d[(size_t)(k - 1) * sxy + (j - jlo) * sx + (i - ilo)]; }
  void fill(double v) { std::fill(d.begin(), d.end(), v); }
};

// profile 1:n (Fortran 1-based), optional 0-based lower bound
struct P1 {
  int lo = 1; std::vector<double> d;
  void alloc(int lo_, int hi_, double v = 0.0) { lo = lo_; d.assign(hi_ - lo_ + 1, v); }
  inline double &operator()(int k) { return d[k - lo]; }
  inline double operator()(int k) const { return d[k - lo]; }
  void fill(double v) { std::fill(d.begin(), d.end(), v); }
};

inline double dmax(double a, double b) { return a > b ? a : b; }
inline double dmin(double a, double b) { return a < b ? a : b; }

// ---- sat.F90 -------------------------------------------------------------------------------------
double esatw_crm(double t);
double qsatw_crm(double t, double p);
double dtesatw_crm(double t);
double dtqsatw_crm(double t, double p);
double esati_crm(double t);
double qsati_crm(double t, double p);
double dtesati_crm(double t);
double dtqsati_crm(double t, double p);

// ---- fft.F90 (real periodic transform in the FFT99 coefficient ordering) -------------------------
struct FftPlan {
  int n = 0;
  std::vector<double> trig;   // cos/sin pairs, length n (complex table for n/2) + n/2.. (half-angle table)
  void init(int n, bool f32_twiddles);
  // a has n+2 entries.  isign=-1: gridpoint -> spectral (a0,b0,a1,b1,...,a_{n/2},b_{n/2}), normalised by 1/n.
  //                      isign=+1: spectral -> gridpoint, unnormalised.   (fft.F90:49-53)
  void fft991(double *a, int isign) const;
};

// ---- RNG_MT.F90 ------------------------------------------------------------------------------------
struct RngMT {
  int mt[624]; int mti = 625;
  void set_seed(int seed);
  double gen_rand();
};

}  // namespace orc
