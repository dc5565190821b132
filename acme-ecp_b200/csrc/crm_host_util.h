// crm_host_util.h -- host-side pieces of the CRM path that the reference also runs on the host:
// Mersenne twister of setperturb (PHYS/RNG_MT.F90:27-78), gamma-function scalars (CRM/gammafff.c:12,
// MICRO_SAM1MOM/precip_init.F90:23-43), FFT trig tables (CRM/fft.F90:247-272 precision semantics).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>
#include "crm_common.cuh"

namespace crm {

// RNG_MT.F90: non-standard UMASK, 32-bit wrap-around seeding, interval (0,1)  
struct RngMT {
  uint32_t mt[624];
  int mti = 625;
  void set_seed(int seed) {
    mt[0] = (uint32_t)seed;
    for (int i = 1; i < 624; i++) mt[i] = 69069u * mt[i - 1];
    mti = 624;
  }
  double gen_rand() {
    const int N = 624, M = 397;
    const uint32_t MATA = 0x9908B0DFu /* -1727483681 */, UMASK = (uint32_t)(-1147483648LL), LMASK = 0x7FFFFFFFu;
    const uint32_t TMASKB = (uint32_t)(-1658038656), TMASKC = (uint32_t)(-272236544);
    if (mti >= N) {
      uint32_t y;
      for (int kk = 0; kk < N - M; kk++) {
        y = (mt[kk] & UMASK) | (mt[kk + 1] & LMASK);
        mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? MATA : 0u);
      }
      for (int kk = N - M; kk < N - 1; kk++) {
        y = (mt[kk] & UMASK) | (mt[kk + 1] & LMASK);
        mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATA : 0u);
      }
      y = (mt[N - 1] & UMASK) | (mt[0] & LMASK);
      mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATA : 0u);
      mti = 0;
    }
    uint32_t y = mt[mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & TMASKB;
    y ^= (y << 15) & TMASKC;
    y ^= (y >> 18);
    const int32_t ys = (int32_t)y;
    if (ys < 0) return ((double)ys + 4294967296.0) / 4294967296.0;
    return ((double)ys + 0.5) / 4294967296.0;
  }
};

static inline float gammafff(float x) { return (float)std::exp(std::lgamma((double)x)); }

// returns false if Gamma(3) != 2 (precip_init.F90:42-45 abort)
static inline bool micro_scalars(MicroScalars &m) {
  const double b_rain = R4(0.8), b_snow = R4(0.25), b_grau = R4(0.5);
  const double a_rain = R4(842.), a_snow = R4(4.84), a_grau = R4(94.5);
  const double rhor = R4(1000.), rhos = R4(100.), rhog = R4(400.);
  const double nzeror = R4(8.e6), nzeros = R4(3.e6), nzerog = R4(4.e6);
  float gam3 = gammafff(3.f);
  float gamr1 = gammafff((float)(3. + b_rain)), gamr2 = gammafff((float)((5. + b_rain) / 2.)), gamr3 = gammafff((float)(4. + b_rain));
  float gams1 = gammafff((float)(3. + b_snow)), gams2 = gammafff((float)((5. + b_snow) / 2.)), gams3 = gammafff((float)(4. + b_snow));
  float gamg1 = gammafff((float)(3. + b_grau)), gamg2 = gammafff((float)((5. + b_grau) / 2.)), gamg3 = gammafff((float)(4. + b_grau));
  if (std::lround(gam3) != 2) return false;
  m.gams1 = gams1; m.gams2 = gams2; m.gamg1 = gamg1; m.gamg2 = gamg2; m.gamr1 = gamr1; m.gamr2 = gamr2;
  m.a_bg = 1. / (CRM_TBGMAX - CRM_TBGMIN);
  m.a_pr = 1. / (CRM_TPRMAX - CRM_TPRMIN);
  m.a_gr = 1. / (CRM_TGRMAX - CRM_TGRMIN);
  m.crain = b_rain / 4.; m.csnow = b_snow / 4.; m.cgrau = b_grau / 4.;
  m.vrain = a_rain * gamr3 / 6. / std::pow(CRM_PI * rhor * nzeror, m.crain);
  m.vsnow = a_snow * gams3 / 6. / std::pow(CRM_PI * rhos * nzeros, m.csnow);
  m.vgrau = a_grau * gamg3 / 6. / std::pow(CRM_PI * rhog * nzerog, m.cgrau);
  return true;
}

// cos/sin(2*pi*q/n), q = 0..n-1.  f32: angle and cos/sin evaluated in REAL(4) like fftrig_crm (implicit typing).
static inline void twiddles(int n, bool f32, std::vector<double> &c, std::vector<double> &s) {
  c.resize(n); s.resize(n);
  if (f32) {
    float pi = 2.0f * asinf(1.0f);
    float del = (float)(((double)pi + (double)pi) / (double)n);
    for (int q = 0; q < n; q++) {
      float angle = (float)((double)q * (double)del);
      c[q] = (double)cosf(angle); s[q] = (double)sinf(angle);
    }
  } else {
    const double pi = 3.14159265358979323846;
    for (int q = 0; q < n; q++) { c[q] = std::cos(2.0 * pi * q / n); s[q] = std::sin(2.0 * pi * q / n); }
  }
}

}  // namespace crm
