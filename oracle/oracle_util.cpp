// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).  Saturation functions, FFT, RNG.
#include "oracle_common.h"
#include <cstdint>

namespace orc {

// ---------------------------------------------------------------------------------------------------
// CRM/sat.F90:12-139.  Flatau et al. polynomials; all coefficients are REAL(4) literals in DATA
// statements assigned to real(crm_rknd) variables.
// ---------------------------------------------------------------------------------------------------
static inline double poly8(const double *a, double dt) {
  return a[0] + dt * (a[1] + dt * (a[2] + dt * (a[3] + dt * (a[4] + dt * (a[5] + dt * (a[6] + dt * (a[7] + a[8] * dt)))))));
}
double esatw_crm(double t) {   // sat.F90:12-32
  static const double a[9] = {R4(6.105851), R4(0.4440316), R4(0.1430341e-1), R4(0.2641412e-3), R4(0.2995057e-5),
                              R4(0.2031998e-7), R4(0.6936113e-10), R4(0.2564861e-13), R4(-0.3704404e-15)};
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(a, dt);
  // 2.*0.01 is a literal-only sub-expression -> evaluated in REAL(4)
  return (double)(2.0f * 0.01f) * std::exp(R4(9.550426) - R4(5723.265) / t + R4(3.53068) * std::log(t) - R4(0.00728332) * t);
}
double qsatw_crm(double t, double p) {   // sat.F90:34-43
  double es = esatw_crm(t);
  return R4(0.622) * es / dmax(es, p - es);
}
double dtesatw_crm(double t) {   // sat.F90:46-65
  static const double a[9] = {R4(0.443956472), R4(0.285976452e-1), R4(0.794747212e-3), R4(0.121167162e-4), R4(0.103167413e-6),
                              R4(0.385208005e-9), R4(-0.604119582e-12), R4(-0.792933209e-14), R4(-0.599634321e-17)};
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(a, dt);
  return esatw_crm(t + 1) - esatw_crm(t);
}
double dtqsatw_crm(double t, double p) { return R4(0.622) * dtesatw_crm(t) / p; }   // sat.F90:67-74
double esati_crm(double t) {   // sat.F90:77-94
  static const double a[9] = {R4(6.11147274), R4(0.503160820), R4(0.188439774e-1), R4(0.420895665e-3), R4(0.615021634e-5),
                              R4(0.602588177e-7), R4(0.385852041e-9), R4(0.146898966e-11), R4(0.252751365e-14)};
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(a, dt);
  return R4(0.01) * std::exp(R4(9.550426) - R4(5723.265) / t + R4(3.53068) * std::log(t) - R4(0.00728332) * t);
}
double qsati_crm(double t, double p) {   // sat.F90:96-105
  double es = esati_crm(t);
  return R4(0.622) * es / dmax(es, p - es);
}
double dtesati_crm(double t) {   // sat.F90:108-127
  static const double a[9] = {R4(0.503223089), R4(0.377174432e-1), R4(0.126710138e-2), R4(0.249065913e-4), R4(0.312668753e-6),
                              R4(0.255653718e-8), R4(0.132073448e-10), R4(0.390204672e-13), R4(0.497275778e-16)};
  double dt = t - R4(273.16);
  if (dt > R4(-80.)) return poly8(a, dt);
  return esati_crm(t + R4(1.)) - esati_crm(t);
}
double dtqsati_crm(double t, double p) { return R4(0.622) * dtesati_crm(t) / p; }   // sat.F90:130-137

// ---------------------------------------------------------------------------------------------------
// CRM/fft.F90: fft991_crm = multiple real/half-complex periodic FFT (Temperton FFT99).  Restated as the
// same mathematical algorithm: half-length complex transform + Cooley/Lewis/Welch pre/post-processing
// (fft.F90:301-466), using trig tables built exactly as fftrig_crm does (fft.F90:247-299), including the
// implicit-REAL(4) pi/del/angle of the reference when f32_twiddles is set.
// The complex passes are radix-2 here (reference: radix 4/2 self-sorting vpassm, fft.F90:466-808), so
// roundoff differs in the last bits; n must be a power of two (configs: 8, 32, 64).
// ---------------------------------------------------------------------------------------------------
void FftPlan::init(int n_, bool f32) {
  n = n_;
  int nh = n / 2;
  trig.assign(2 * nh + 2 * (nh / 2 + 1), 0.0);
  if (f32) {
    float pi = 2.0f * asinf(1.0f);                                  // fft.F90:251
    float del = (float)(((double)pi + (double)pi) / (double)nh);    // :256
    for (int k = 0; k < nh; k++) {
      float angle = (float)(0.5 * (double)(2 * k) * (double)del);   // :259
      trig[2 * k] = (double)cosf(angle);
      trig[2 * k + 1] = (double)sinf(angle);
    }
    float del2 = (float)(0.5 * (double)del);                        // :265
    for (int m = 0; m <= nh / 2; m++) {
      float angle = (float)(0.5 * (double)(2 * m) * (double)del2);  // :270
      trig[2 * nh + 2 * m] = (double)cosf(angle);
      trig[2 * nh + 2 * m + 1] = (double)sinf(angle);
    }
  } else {
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < nh; k++) {
      trig[2 * k] = std::cos(2.0 * pi * k / nh);
      trig[2 * k + 1] = std::sin(2.0 * pi * k / nh);
    }
    for (int m = 0; m <= nh / 2; m++) {
      trig[2 * nh + 2 * m] = std::cos(2.0 * pi * m / n);
      trig[2 * nh + 2 * m + 1] = std::sin(2.0 * pi * m / n);
    }
  }
}

// in-place complex FFT, length nh (power of 2):  Z_k = sum_m z_m exp(sign*2*pi*i*m*k/nh)
static void cfft(int nh, double *re, double *im, int sign, const double *T) {
  for (int i = 1, j = 0; i < nh; i++) {
    int bit = nh >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
  }
  for (int len = 2; len <= nh; len <<= 1) {
    int half = len / 2, step = nh / len;
    for (int i = 0; i < nh; i += len)
      for (int j = 0; j < half; j++) {
        double wr = T[2 * j * step], wi = sign * T[2 * j * step + 1];
        double xr = re[i + j + half], xi = im[i + j + half];
        double vr = wr * xr - wi * xi, vi = wr * xi + wi * xr;
        double ur = re[i + j], ui = im[i + j];
        re[i + j] = ur + vr; im[i + j] = ui + vi;
        re[i + j + half] = ur - vr; im[i + j + half] = ui - vi;
      }
  }
}

void FftPlan::fft991(double *a, int isign) const {
  const int nh = n / 2;
  const double *T = trig.data();
  const double *H = trig.data() + 2 * nh;   // H[2m],H[2m+1] = cos,sin(2 pi m / n), m = 0..nh/2
  std::vector<double> re(nh), im(nh);
  if (isign == -1) {
    for (int m = 0; m < nh; m++) { re[m] = a[2 * m]; im[m] = a[2 * m + 1]; }
    cfft(nh, re.data(), im.data(), -1, T);
    double scale = 1.0 / (double)n;
    double a0 = scale * (re[0] + im[0]), an = scale * (re[0] - im[0]);
    std::vector<double> out(n + 2, 0.0);
    out[0] = a0; out[1] = 0.0; out[n] = an; out[n + 1] = 0.0;
    scale = 0.5 * scale;
    for (int m = 1; 2 * m < nh; m++) {          // pairs (m, nh-m)   (fft.F90:404-430)
      int b = nh - m;
      double c = H[2 * m], s = H[2 * m + 1];
      double sr = re[m] + re[b], dr = re[m] - re[b], si = im[m] + im[b], di = im[m] - im[b];
      // X_m = E + exp(-i th) O ;  X_b = conj(E) - exp(+i th) conj(O)
      out[2 * m]     = scale * (sr + (c * si - s * dr));
      out[2 * m + 1] = scale * (di - (c * dr + s * si));
      out[2 * b]     = scale * (sr - (c * si - s * dr));
      out[2 * b + 1] = scale * (-di - (c * dr + s * si));
    }
    if (nh >= 2) {                               // wavenumber n/4  (fft.F90:432-444)
      int m = nh / 2;
      out[2 * m] = 2.0 * scale * re[m];
      out[2 * m + 1] = -2.0 * scale * im[m];
    }
    for (int i = 0; i < n + 2; i++) a[i] = out[i];
  } else {
    // spectral -> gridpoint; b(0) and b(n/2) are ignored like fft99a (fft.F90:313-320)
    re[0] = a[0] + a[n];
    im[0] = a[0] - a[n];
    for (int m = 1; 2 * m < nh; m++) {
      int b = nh - m;
      double c = H[2 * m], s = H[2 * m + 1];
      double am = a[2 * m], bm = a[2 * m + 1], ab = a[2 * b], bb = a[2 * b + 1];
      // Ze = c_m + conj(c_b), Zo = (c_m - conj(c_b)) * exp(i th);  Z_m = Ze + i Zo
      double er = am + ab, ei = bm - bb, dr = am - ab, di = bm + bb;
      double zor = c * dr - s * di, zoi = c * di + s * dr;
      re[m] = er - zoi; im[m] = ei + zor;
      // for b = nh-m: th_b = pi - th -> exp(i th_b) = -c + i s ; c_b - conj(c_m) = (-dr) + i(di)
      double zor2 = (-c) * (-dr) - s * di, zoi2 = (-c) * di + s * (-dr);
      re[b] = er - zoi2; im[b] = -ei + zor2;
    }
    if (nh >= 2) {
      int m = nh / 2;
      re[m] = 2.0 * a[2 * m]; im[m] = -2.0 * a[2 * m + 1];
    }
    cfft(nh, re.data(), im.data(), +1, T);
    for (int m = 0; m < nh; m++) { a[2 * m] = re[m]; a[2 * m + 1] = im[m]; }
    a[n] = 0.0; a[n + 1] = 0.0;                  // fft.F90:139-147
  }
}

// ---------------------------------------------------------------------------------------------------
// components/cam/src/physics/cam/RNG_MT.F90:27-78 (Mersenne twister with the reference's non-standard
// UMASK and 32-bit wrap-around seeding)
// ---------------------------------------------------------------------------------------------------
void RngMT::set_seed(int seed) {
  for (int i = 0; i < 624; i++) mt[i] = 625;
  mt[0] = seed;
  for (int i = 1; i < 624; i++) mt[i] = (int)((uint32_t)69069u * (uint32_t)mt[i - 1]);
  mti = 624;   // value of the DO variable after loop completion
}
double RngMT::gen_rand() {
  const int N = 624, M = 397;
  const uint32_t MATA = (uint32_t)(-1727483681), UMASK = (uint32_t)(-1147483648LL), LMASK = 2147483647u;
  const uint32_t TMASKB = (uint32_t)(-1658038656), TMASKC = (uint32_t)(-272236544);
  const uint32_t mag01[2] = {0u, MATA};
  if (mti >= N) {
    if (mti == N + 1) { set_seed(4357); for (int i = 0; i < 624; i++) mt[i] = 625; }
    uint32_t *m = (uint32_t *)mt;
    uint32_t y;
    for (int kk = 0; kk <= N - M - 1; kk++) {
      y = (m[kk] & UMASK) | (m[kk + 1] & LMASK);
      m[kk] = m[kk + M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (int kk = N - M; kk <= N - 2; kk++) {
      y = (m[kk] & UMASK) | (m[kk + 1] & LMASK);
      m[kk] = m[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    y = (m[N - 1] & UMASK) | (m[0] & LMASK);
    m[N - 1] = m[M - 1] ^ (y >> 1) ^ mag01[y & 1u];
    mti = 0;
  }
  uint32_t y = (uint32_t)mt[mti];
  mti++;
  y ^= (y >> 11);
  y ^= (y << 7) & TMASKB;
  y ^= (y << 15) & TMASKC;
  y ^= (y >> 18);
  int ys = (int)y;
  if (ys < 0) return ((double)ys + 4294967296.0) / 4294967296.0;
  return ((double)ys + 0.5) / 4294967296.0;
}

}  // namespace orc
