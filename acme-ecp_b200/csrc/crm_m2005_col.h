// crm_m2005_col.h -- Morrison et al. (2005) two-moment microphysics for ONE CRM column: restatement of
//   GRAUPEL_INIT            reference CRM/MICRO_M2005/module_mp_graupel.F90:346-758
//   M2005MICRO_GRAUPEL      module_mp_graupel.F90:1102-5610   (build: CRM, not CLUBB_CRM / ECPP / MODAL_AERO)
//   POLYSVP, GAMMA, DERF1   module_mp_graupel.F90:5614, 5685, 5919  (REAL functions: results rounded to single)
//   satadj_liquid           CRM/MICRO_M2005/microphysics.F90:1435-1500
// with the options micro_setparm fixes (microphysics.F90:288-469): doicemicro, dograupel, dopredictNc, dosubgridw,
// docloudedgeactivation = .true.; dohail, dosb_warm_rain, doarcticicenucl, dofix_pgam = .false.; aerosol_mode = 2
// => INUM 0, IACT 2, IBASE 2, ISUB 0, ILIQ 0, INUC 0, IGRAUP 0, IHAIL 0, IRAIN 0, ISATADJ 1.  Branches of other option values are
// not restated.  Literals without a kind suffix are REAL(4) in the reference build (no -r8): F4() reproduces them, and
// literal-only sub-expressions are evaluated in single precision.
//
// Plain C++ (no CUDA types): compiled by nvcc into the product kernel k_m2005_column (crm_k_m2005.cuh) and by g++ into the
// CPU oracle.  This ONE transcription is shared by both - the 2 800-line column physics is not restated twice - so GPU-vs-
// oracle parity of cfg 3 checks compilers, memory plumbing and the driver around it, not the transcription; the
// transcription itself is pinned by the known-answer tests in tests/test_m2005_kat.py.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define M25_HD __host__ __device__
#define M25_NI __noinline__   // the big helpers are called out of line: the column kernel's code size decides its instruction-cache behaviour
#else
#define M25_HD
#define M25_NI
#endif
#define F4(x) ((double)(x##f))
// x**y of the scheme (positive bases, real exponents).  Host builds (oracle, emulator): libm pow.  Device: ONE out-of-line copy of
// exp(y*log(x)) instead of 125 inlined copies of CUDA's pow (each ~250 instructions: 52 % of the kernel's instructions and an
// instruction-cache footprint that stalled the schedulers 31 % of the time); relative error <= (|y ln x| + 2) * 2^-53 ~ 1e-14.
#if defined(__CUDA_ARCH__)
__device__ __noinline__ double m25_pow_dev(double x, double y) {
  if (y == 0.) return 1.;
  return exp(y * log(x));
}
__device__ __noinline__ double m25_exp_dev(double x) { return exp(x); }
__device__ __noinline__ double m25_log_dev(double x) { return log(x); }
#define M25_POW(x, y) m25_pow_dev(x, y)
// the warps of a block start every level of the process-rate pass together: they then fetch the same instructions at about the same
// time (the pass is bound by instruction-cache misses); the launch guarantees that every thread of the block runs the loop
#define M25_LEVEL_SYNC() __syncthreads()
#define M25_EXP(x) m25_exp_dev(x)
#define M25_LOG(x) m25_log_dev(x)
#else
#define M25_POW(x, y) pow(x, y)
#define M25_LEVEL_SYNC()
#define M25_EXP(x) exp(x)
#define M25_LOG(x) log(x)
#endif

namespace m2005 {

// ---- array ids of the per-column work space (A::operator()(id, k), k = 1..nz) ----
enum {
  QC = 0, QI, QNI, QR, NC, NI, NS, NR, QG, NG, T3, QV,                          // state of the column (in/out)
  TQC, TQI, TQNI, TQR, TNC, TNI, TNS, TNR, TQG, TNG, TT, TQV,                   // tendencies *3DTEN
  SQR, SQI, SQNI, SQC, SQG,                                                     // sedimentation tendencies Q*STEN
  DR, DI_, DNI, DS_, DNS, DNR, DC, DNC, DG_, DNG,                               // DUMR, DUMI, DUMFNI, DUMQS, DUMFNS, DUMFNR, DUMC, DUMFNC, DUMG, DUMFNG
  FR, FI, FNI, FS, FNS, FNR, FC, FNC, FG, FNG,                                  // fall speeds
  LAMI_, AIN_, ARN_, ASN_, ACN_, AGN_,
  AUT, ACC, EVPC, EVPR, MLT, SUB, DEP, CON,                                     // process-rate outputs (aut1d ...)
  PRES, RHO, DZQ, W3, WVAR,                                                     // inputs
  NARR
};

struct Const {   // GRAUPEL_INIT (module_mp_graupel.F90:346-758)
  double PI, SQRTPI;
  double NDCNST;
  double AI, AC, AS, AR, AG, BI, BC, BS, BR, BG;
  double R, RV, CP, XXLV, XXLS, EP_2;
  double TMELT, RHOSU, RHOW, RHOI, RHOSN, RHOG, AIMM, BIMM, ECR, DCS, MI0, MG0, F1S, F2S, F1R, F2R, G, QSMALL, EII, ECI, CPW;
  double CI, DI, CS, DS, CG, DG, RIN, MMULT;
  double LAMMAXI, LAMMINI, LAMMAXR, LAMMINR, LAMMAXS, LAMMINS, LAMMAXG, LAMMING;
  double MW, OSM, VI, EPSM, RHOA, MAP, MA, RR, BACT, RM1, RM2, NANEW1, NANEW2, SIG1, SIG2, F11, F12, F21, F22;
  double CONS[42];   // CONS1..CONS41 (index = number)
  // satadj_liquid / micro_proc constants (params.F90)
  double fac_cond, fac_sub, lcond;
};

// POLYSVP (module_mp_graupel.F90:5614-5681): Flatau et al. polynomials (coefficients REAL(4) literals), result REAL(4)
M25_HD M25_NI inline double polysvp(double T, int type) {
  double dt = T - F4(273.16);
  if (dt < F4(-80.)) dt = F4(-80.);
  double p;
  if (type == 1)
    p = F4(6.11147274) + dt * (F4(0.503160820) + dt * (F4(0.188439774e-1) + dt * (F4(0.420895665e-3) + dt * (F4(0.615021634e-5) + dt * (F4(0.602588177e-7) +
        dt * (F4(0.385852041e-9) + dt * (F4(0.146898966e-11) + F4(0.252751365e-14) * dt)))))));
  else
    p = F4(6.11239921) + dt * (F4(0.443987641) + dt * (F4(0.142986287e-1) + dt * (F4(0.264847430e-3) + dt * (F4(0.302950461e-5) + dt * (F4(0.206739458e-7) +
        dt * (F4(0.640689451e-10) + dt * (F4(-0.952447341e-13) + F4(-0.976195544e-15) * dt)))))));
  p = p * F4(100.);
  return (double)(float)p;
}

// GAMMA (module_mp_graupel.F90:5685-5916): W. J. Cody's algorithm with REAL(4) DATA constants, result REAL(4)
M25_HD M25_NI inline double gamma_f(double X, double PI, double SQRTPI) {
  const double ONE = 1., HALF = 0.5, TWELVE = 12., TWO = 2., ZERO = 0.;
  const double XBIG = F4(35.040E0), XMININ = F4(1.18E-38), EPS = F4(1.19E-7), XINF = F4(3.4E38);
  const double P[8] = {F4(-1.71618513886549492533811E+0), F4(2.47656508055759199108314E+1), F4(-3.79804256470945635097577E+2),
                       F4(6.29331155312818442661052E+2), F4(8.66966202790413211295064E+2), F4(-3.14512729688483675254357E+4),
                       F4(-3.61444134186911729807069E+4), F4(6.64561438202405440627855E+4)};
  const double Q[8] = {F4(-3.08402300119738975254353E+1), F4(3.15350626979604161529144E+2), F4(-1.01515636749021914166146E+3),
                       F4(-3.10777167157231109440444E+3), F4(2.25381184209801510330112E+4), F4(4.75584627752788110767815E+3),
                       F4(-1.34659959864969306392456E+5), F4(-1.15132259675553483497211E+5)};
  const double C[7] = {F4(-1.910444077728E-03), F4(8.4171387781295E-04), F4(-5.952379913043012E-04), F4(7.93650793500350248E-04),
                       F4(-2.777777777777681622553E-03), F4(8.333333333333333331554247E-02), F4(5.7083835261E-03)};
  bool parity = false;
  double FACT = ONE, RES, Y = X, Y1, Z;
  int N = 0;
  if (Y <= ZERO) {
    Y = -X;
    Y1 = trunc(Y);
    RES = Y - Y1;
    if (RES != ZERO) {
      if (Y1 != trunc(Y1 * HALF) * TWO) parity = true;
      FACT = -PI / sin(PI * RES);
      Y = Y + ONE;
    } else return (double)(float)XINF;
  }
  if (Y < EPS) {
    if (Y >= XMININ) RES = ONE / Y;
    else return (double)(float)XINF;
  } else if (Y < TWELVE) {
    Y1 = Y;
    if (Y < ONE) { Z = Y; Y = Y + ONE; }
    else { N = (int)Y - 1; Y = Y - (double)(float)N; Z = Y - ONE; }
    double XNUM = ZERO, XDEN = ONE;
    for (int I = 0; I < 8; I++) { XNUM = (XNUM + P[I]) * Z; XDEN = XDEN * Z + Q[I]; }
    RES = XNUM / XDEN + ONE;
    if (Y1 < Y) RES = RES / Y1;
    else if (Y1 > Y) { for (int I = 1; I <= N; I++) { RES = RES * Y; Y = Y + ONE; } }
  } else {
    if (Y <= XBIG) {
      const double YSQ = Y * Y;
      double SUM = C[6];
      for (int I = 0; I < 6; I++) SUM = SUM / YSQ + C[I];
      SUM = SUM / Y - Y + SQRTPI;
      SUM = SUM + (Y - HALF) * M25_LOG(Y);
      RES = M25_EXP(SUM);
    } else return (double)(float)XINF;
  }
  if (parity) RES = -RES;
  if (FACT != ONE) RES = FACT / RES;
  return (double)(float)RES;
}

// DERF1 (module_mp_graupel.F90:5919-6030), REAL(4) DATA coefficients, result REAL(4)
M25_HD M25_NI inline double derf1(double X) {
  static constexpr double A[65] = {
      F4(0.00000000005958930743E0), F4(-0.00000000113739022964E0), F4(0.00000001466005199839E0), F4(-0.00000016350354461960E0),
      F4(0.00000164610044809620E0), F4(-0.00001492559551950604E0), F4(0.00012055331122299265E0), F4(-0.00085483269811296660E0),
      F4(0.00522397762482322257E0), F4(-0.02686617064507733420E0), F4(0.11283791670954881569E0), F4(-0.37612638903183748117E0),
      F4(1.12837916709551257377E0),
      F4(0.00000000002372510631E0), F4(-0.00000000045493253732E0), F4(0.00000000590362766598E0), F4(-0.00000006642090827576E0),
      F4(0.00000067595634268133E0), F4(-0.00000621188515924000E0), F4(0.00005103883009709690E0), F4(-0.00037015410692956173E0),
      F4(0.00233307631218880978E0), F4(-0.01254988477182192210E0), F4(0.05657061146827041994E0), F4(-0.21379664776456006580E0),
      F4(0.84270079294971486929E0),
      F4(0.00000000000949905026E0), F4(-0.00000000018310229805E0), F4(0.00000000239463074000E0), F4(-0.00000002721444369609E0),
      F4(0.00000028045522331686E0), F4(-0.00000261830022482897E0), F4(0.00002195455056768781E0), F4(-0.00016358986921372656E0),
      F4(0.00107052153564110318E0), F4(-0.00608284718113590151E0), F4(0.02986978465246258244E0), F4(-0.13055593046562267625E0),
      F4(0.67493323603965504676E0),
      F4(0.00000000000382722073E0), F4(-0.00000000007421598602E0), F4(0.00000000097930574080E0), F4(-0.00000001126008898854E0),
      F4(0.00000011775134830784E0), F4(-0.00000111992758382650E0), F4(0.00000962023443095201E0), F4(-0.00007404402135070773E0),
      F4(0.00050689993654144881E0), F4(-0.00307553051439272889E0), F4(0.01668977892553165586E0), F4(-0.08548534594781312114E0),
      F4(0.56909076642393639985E0),
      F4(0.00000000000155296588E0), F4(-0.00000000003032205868E0), F4(0.00000000040424830707E0), F4(-0.00000000471135111493E0),
      F4(0.00000005011915876293E0), F4(-0.00000048722516178974E0), F4(0.00000430683284629395E0), F4(-0.00003445026145385764E0),
      F4(0.00024879276133931664E0), F4(-0.00162940941748079288E0), F4(0.00988786373932350462E0), F4(-0.05962426839442303805E0),
      F4(0.49766113250947636708E0)};
  static constexpr double B[65] = {
      F4(-0.00000000029734388465E0), F4(0.00000000269776334046E0), F4(-0.00000000640788827665E0), F4(-0.00000001667820132100E0),
      F4(-0.00000021854388148686E0), F4(0.00000266246030457984E0), F4(0.00001612722157047886E0), F4(-0.00025616361025506629E0),
      F4(0.00015380842432375365E0), F4(0.00815533022524927908E0), F4(-0.01402283663896319337E0), F4(-0.19746892495383021487E0),
      F4(0.71511720328842845913E0),
      F4(-0.00000000001951073787E0), F4(-0.00000000032302692214E0), F4(0.00000000522461866919E0), F4(0.00000000342940918551E0),
      F4(-0.00000035772874310272E0), F4(0.00000019999935792654E0), F4(0.00002687044575042908E0), F4(-0.00011843240273775776E0),
      F4(-0.00080991728956032271E0), F4(0.00661062970502241174E0), F4(0.00909530922354827295E0), F4(-0.20160072778491013140E0),
      F4(0.51169696718727644908E0),
      F4(0.00000000003147682272E0), F4(-0.00000000048465972408E0), F4(0.00000000063675740242E0), F4(0.00000003377623323271E0),
      F4(-0.00000015451139637086E0), F4(-0.00000203340624738438E0), F4(0.00001947204525295057E0), F4(0.00002854147231653228E0),
      F4(-0.00101565063152200272E0), F4(0.00271187003520095655E0), F4(0.02328095035422810727E0), F4(-0.16725021123116877197E0),
      F4(0.32490054966649436974E0),
      F4(0.00000000002319363370E0), F4(-0.00000000006303206648E0), F4(-0.00000000264888267434E0), F4(0.00000002050708040581E0),
      F4(0.00000011371857327578E0), F4(-0.00000211211337219663E0), F4(0.00000368797328322935E0), F4(0.00009823686253424796E0),
      F4(-0.00065860243990455368E0), F4(-0.00075285814895230877E0), F4(0.02585434424202960464E0), F4(-0.11637092784486193258E0),
      F4(0.18267336775296612024E0),
      F4(-0.00000000000367789363E0), F4(0.00000000020876046746E0), F4(-0.00000000193319027226E0), F4(-0.00000000435953392472E0),
      F4(0.00000018006992266137E0), F4(-0.00000078441223763969E0), F4(-0.00000675407647949153E0), F4(0.00008428418334440096E0),
      F4(-0.00017604388937031815E0), F4(-0.00239729611435071610E0), F4(0.02064129023876022970E0), F4(-0.06905562880005864105E0),
      F4(0.09084526782065478489E0)};
  const double W = fabs(X);
  double Y;
  if (W < 2.2) {
    double T = W * W;
    int K = (int)T;
    T = T - K;
    K = K * 13;
    Y = ((((((((((((A[K] * T + A[K + 1]) * T + A[K + 2]) * T + A[K + 3]) * T + A[K + 4]) * T + A[K + 5]) * T + A[K + 6]) * T + A[K + 7]) * T +
            A[K + 8]) * T + A[K + 9]) * T + A[K + 10]) * T + A[K + 11]) * T + A[K + 12]) * W;
  } else if (W < 6.9) {
    int K = (int)W;
    const double T = W - K;
    K = 13 * (K - 2);
    Y = (((((((((((B[K] * T + B[K + 1]) * T + B[K + 2]) * T + B[K + 3]) * T + B[K + 4]) * T + B[K + 5]) * T + B[K + 6]) * T + B[K + 7]) * T +
           B[K + 8]) * T + B[K + 9]) * T + B[K + 10]) * T + B[K + 11]) * T + B[K + 12];
    Y = Y * Y; Y = Y * Y; Y = Y * Y;
    Y = 1 - Y * Y;
  } else Y = 1;
  if (X < 0) Y = -Y;
  return (double)(float)Y;
}

// GRAUPEL_INIT.  rgas, rv, cp, lcond, lsub: CRM/params.F90 constants (doubles).  Host only.
inline void graupel_init(Const &c, double rgas, double rv, double cp, double lcond, double lsub) {
  c.PI = F4(3.1415926535897932384626434);
  c.SQRTPI = F4(0.9189385332046727417803297);
  auto G_ = [&](double x) { return gamma_f(x, c.PI, c.SQRTPI); };
  c.NDCNST = F4(100.);   // Nc0
  c.AI = F4(700.); c.AC = F4(3.E7); c.AS = F4(11.72); c.AR = F4(841.99667);
  c.BI = F4(1.); c.BC = F4(2.); c.BS = F4(0.41); c.BR = F4(0.8);
  c.AG = F4(19.3); c.BG = F4(0.37);
  c.R = rgas; c.RV = rv; c.CP = cp; c.XXLV = lcond; c.XXLS = lsub; c.EP_2 = rgas / rv;
  c.TMELT = F4(273.15);
  c.RHOSU = F4(85000.) / (c.R * c.TMELT);
  c.RHOW = F4(997.); c.RHOI = F4(500.); c.RHOSN = F4(100.); c.RHOG = F4(400.);
  c.AIMM = F4(0.66); c.BIMM = F4(100.); c.ECR = F4(1.);
  c.DCS = F4(125.E-6);
  { const float e = 10.E-6f; c.MI0 = (double)(4.f / 3.f) * c.PI * c.RHOI * (double)(e * e * e); }
  c.MG0 = F4(1.6E-10);
  c.F1S = F4(0.86); c.F2S = F4(0.28); c.F1R = F4(0.78); c.F2R = F4(0.308);
  c.G = F4(9.806); c.QSMALL = F4(1.E-14); c.EII = F4(0.1); c.ECI = F4(0.7); c.CPW = F4(4218.);
  c.CI = c.RHOI * c.PI / F4(6.); c.DI = F4(3.);
  c.CS = c.RHOSN * c.PI / F4(6.); c.DS = F4(3.);
  c.CG = c.RHOG * c.PI / F4(6.); c.DG = F4(3.);
  c.RIN = F4(0.1E-6);
  { const float e = 5.E-6f; c.MMULT = (double)(4.f / 3.f) * c.PI * c.RHOI * (double)(e * e * e); }
  c.LAMMAXI = (double)(1.f / 1.E-6f);
  c.LAMMINI = F4(1.) / (F4(2.) * c.DCS + F4(100.E-6));
  c.LAMMAXR = (double)(1.f / 20.E-6f);
  c.LAMMINR = (double)(1.f / 2800.E-6f);
  c.LAMMAXS = (double)(1.f / 10.E-6f);
  c.LAMMINS = (double)(1.f / 2000.E-6f);
  c.LAMMAXG = (double)(1.f / 20.E-6f);
  c.LAMMING = (double)(1.f / 2000.E-6f);
  c.MW = F4(0.018); c.OSM = F4(1.); c.VI = F4(3.); c.EPSM = F4(0.7); c.RHOA = F4(1777.); c.MAP = F4(0.132); c.MA = F4(0.0284); c.RR = F4(8.3187);
  c.BACT = c.VI * c.OSM * c.EPSM * c.MW * c.RHOA / (c.MAP * c.RHOW);
  // micro_setparm (microphysics.F90:412-440): two log-normal aerosol modes, um -> m, cm-3 -> m-3
  c.RM1 = F4(1.e-6) * F4(0.052); c.SIG1 = F4(2.04); c.NANEW1 = F4(1.e6) * (double)2500;
  c.RM2 = F4(1.e-6) * F4(1.3); c.SIG2 = F4(2.5); c.NANEW2 = F4(1.e6) * F4(1.8);
  c.F11 = F4(0.5) * M25_EXP(F4(2.5) * (M25_LOG(c.SIG1) * M25_LOG(c.SIG1)));
  c.F21 = F4(1.) + F4(0.25) * M25_LOG(c.SIG1);
  c.F12 = F4(0.5) * M25_EXP(F4(2.5) * (M25_LOG(c.SIG2) * M25_LOG(c.SIG2)));
  c.F22 = F4(1.) + F4(0.25) * M25_LOG(c.SIG2);
  double *C = c.CONS;
  C[0] = 0.;
  C[1] = G_(F4(1.) + c.DS) * c.CS;
  C[2] = G_(F4(1.) + c.DG) * c.CG;
  C[3] = G_(F4(4.) + c.BS) / F4(6.);
  C[4] = G_(F4(4.) + c.BR) / F4(6.);
  C[5] = G_(F4(1.) + c.BS);
  C[6] = G_(F4(1.) + c.BR);
  C[7] = G_(F4(4.) + c.BG) / F4(6.);
  C[8] = G_(F4(1.) + c.BG);
  C[9] = G_((double)(5.f / 2.f) + c.BR / F4(2.));
  C[10] = G_((double)(5.f / 2.f) + c.BS / F4(2.));
  C[11] = G_((double)(5.f / 2.f) + c.BG / F4(2.));
  C[12] = G_(F4(1.) + c.DI) * c.CI;
  C[13] = G_(c.BS + F4(3.)) * c.PI / F4(4.) * c.ECI;
  C[14] = G_(c.BG + F4(3.)) * c.PI / F4(4.) * c.ECI;
  C[15] = F4(-1108.) * c.EII * M25_POW(c.PI, (F4(1.) - c.BS) / F4(3.)) * M25_POW(c.RHOSN, (F4(-2.) - c.BS) / F4(3.)) / (double)(4.f * 720.f);
  C[16] = G_(c.BI + F4(3.)) * c.PI / F4(4.) * c.ECI;
  C[17] = (double)(4.f * 2.f * 3.f) * c.RHOSU * c.PI * c.ECI * c.ECI * G_(F4(2.) * c.BS + F4(2.)) / (F4(8.) * (c.RHOG - c.RHOSN));
  C[18] = c.RHOSN * c.RHOSN;
  C[19] = c.RHOW * c.RHOW;
  C[20] = F4(20.) * c.PI * c.PI * c.RHOW * c.BIMM;
  C[21] = F4(4.) / (c.DCS * c.RHOI);
  C[22] = c.PI * c.RHOI * (c.DCS * c.DCS * c.DCS) / F4(6.);
  C[23] = c.PI / F4(4.) * c.EII * G_(c.BS + F4(3.));
  C[24] = c.PI / F4(4.) * c.ECR * G_(c.BR + F4(3.));
  C[25] = c.PI * c.PI / F4(24.) * c.RHOW * c.ECR * G_(c.BR + F4(6.));
  C[26] = c.PI / F4(6.) * c.RHOW;
  C[27] = G_(F4(1.) + c.BI);
  C[28] = G_(F4(4.) + c.BI) / F4(6.);
  { const float e = 25.E-6f; C[29] = (double)(4.f / 3.f) * c.PI * c.RHOW * (double)(e * e * e); }
  C[30] = (double)(4.f / 3.f) * c.PI * c.RHOW;
  C[31] = c.PI * c.PI * c.ECR * c.RHOSN;
  C[32] = c.PI / F4(2.) * c.ECR;
  C[33] = c.PI * c.PI * c.ECR * c.RHOG;
  C[34] = (double)(5.f / 2.f) + c.BR / F4(2.);
  C[35] = (double)(5.f / 2.f) + c.BS / F4(2.);
  C[36] = (double)(5.f / 2.f) + c.BG / F4(2.);
  C[37] = F4(4.) * c.PI * F4(1.38E-23) / (F4(6.) * c.PI * c.RIN);
  C[38] = c.PI * c.PI / F4(3.) * c.RHOW;
  C[39] = c.PI * c.PI / F4(36.) * c.RHOW * c.BIMM;
  C[40] = c.PI / F4(6.) * c.BIMM;
  C[41] = c.PI * c.PI * c.ECR * c.RHOW;
  c.fac_cond = lcond / cp; c.fac_sub = lsub / cp; c.lcond = lcond;
}

// satadj_liquid for one level (microphysics.F90:1435-1500); qt in: total water, out: vapour
M25_HD inline void satadj_liquid(const Const &c, double &tabs, double &qt, double &qc, double pres) {
  double tabs1 = tabs;
  double esat1 = polysvp(tabs1, 0);
  double qsat1 = F4(0.622) * esat1 / (pres - esat1);
  qc = 0.;
  if (qt > qsat1) {
    double d0 = qt - qsat1; if (d0 < 0.) d0 = 0.;
    double dtabs = +c.fac_cond * d0 / (F4(1.) + c.lcond * c.lcond * qsat1 / (c.CP * c.RV * (tabs1 * tabs1)));
    tabs1 = tabs1 + dtabs;
    int niter = 1;
    double thresh = F4(0.01) * c.fac_cond * qsat1;
    if (F4(0.01) < thresh) thresh = F4(0.01);
    while (fabs(dtabs) > thresh && niter < 20) {
      esat1 = polysvp(tabs1, 0);
      qsat1 = F4(0.622) * esat1 / (pres - esat1);
      double d1 = qt - qsat1; if (d1 < 0.) d1 = 0.;
      const double fff = tabs - tabs1 + c.fac_cond * d1;
      const double dfff = F4(1.) + c.lcond * c.lcond * qsat1 / (c.CP * c.RV * (tabs1 * tabs1));
      dtabs = fff / dfff;
      tabs1 = tabs1 + dtabs;
      niter = niter + 1;
    }
    double dq = tabs1 - tabs; if (dq < 0.) dq = 0.;
    qc = dq / c.fac_cond;
    qt = qt - qc;
    tabs = tabs1;
  }
}

M25_HD inline double mx_(double a, double b) { return a > b ? a : b; }
M25_HD inline double mn_(double a, double b) { return a < b ? a : b; }
M25_HD inline double p2_(double x) { return x * x; }
M25_HD inline double p3_(double x) { return x * x * x; }
M25_HD inline double p4_(double x) { const double y = x * x; return y * y; }

// droplet activation on the two log-normal aerosol modes, IACT = 2 (module_mp_graupel.F90:2789-2819 == 4580-4610).
// DUM: vertical velocity (m/s).  Returns the activated number mixing ratio limited by the aerosol number.
template <class Cn>
M25_HD inline double activate2(const Cn &c, double DUM, double T, double EVS, double DV, double KAP, double PRES_, double RHO_) {
  const double SIGVL = F4(0.0761) - F4(1.55E-4) * (T - c.TMELT);
  const double AACT = F4(2.) * c.MW / (c.RHOW * c.RR) * SIGVL / T;
  const double ALPHA = c.G * c.MW * c.XXLV / (c.CP * c.RR * (T * T)) - c.G * c.MA / (c.RR * T);
  const double GAMM = c.RR * T / (EVS * c.MW) + c.MW * (c.XXLV * c.XXLV) / (c.CP * PRES_ * c.MA * T);
  const double GG = F4(1.) / (c.RHOW * c.RR * T / (EVS * DV * c.MW) + c.XXLV * c.RHOW / (KAP * T) * (c.XXLV * c.MW / (T * c.RR) - F4(1.)));
  const double PSI = (double)(2.f / 3.f) * M25_POW(ALPHA * DUM / GG, F4(0.5)) * AACT;
  const double ETA1 = M25_POW(ALPHA * DUM / GG, F4(1.5)) / (F4(2.) * c.PI * c.RHOW * GAMM * c.NANEW1);
  const double ETA2 = M25_POW(ALPHA * DUM / GG, F4(1.5)) / (F4(2.) * c.PI * c.RHOW * GAMM * c.NANEW2);
  const double SM1 = F4(2.) / M25_POW(c.BACT, F4(0.5)) * M25_POW(AACT / (F4(3.) * c.RM1), F4(1.5));
  const double SM2 = F4(2.) / M25_POW(c.BACT, F4(0.5)) * M25_POW(AACT / (F4(3.) * c.RM2), F4(1.5));
  double DUM1 = F4(1.) / (SM1 * SM1) * (c.F11 * M25_POW(PSI / ETA1, F4(1.5)) + c.F21 * M25_POW((SM1 * SM1) / (ETA1 + F4(3.) * PSI), F4(0.75)));
  double DUM2 = F4(1.) / (SM2 * SM2) * (c.F12 * M25_POW(PSI / ETA2, F4(1.5)) + c.F22 * M25_POW((SM2 * SM2) / (ETA2 + F4(3.) * PSI), F4(0.75)));
  const double SMAX = F4(1.) / M25_POW(DUM1 + DUM2, F4(0.5));
  const double UU1 = F4(2.) * M25_LOG(SM1 / SMAX) / (F4(4.242) * M25_LOG(c.SIG1));
  const double UU2 = F4(2.) * M25_LOG(SM2 / SMAX) / (F4(4.242) * M25_LOG(c.SIG2));
  DUM1 = c.NANEW1 / F4(2.) * (F4(1.) - derf1(UU1));
  DUM2 = c.NANEW2 / F4(2.) * (F4(1.) - derf1(UU2));
  DUM2 = (DUM1 + DUM2) / RHO_;
  DUM2 = mn_((c.NANEW1 + c.NANEW2) / RHO_, DUM2);
  return DUM2;
}

// M2005MICRO_GRAUPEL for one column of nz levels (k = 1 bottom).  `a(id, k)` is the work space (ids above); on entry the state
// (QC..NG, T3, QV) and the inputs (PRES [Pa], RHO, DZQ, W3, WVAR) are filled, on return the state is updated, the sedimentation
// tendencies SQ* and the rate outputs AUT..CON are set.  PRECRT / SNOWRT: surface precipitation of the step (mm).
// PH = 0: the whole scheme (oracle, emulator).  PH = 1, 2, 3: the three sections as separate passes over the work space (process rates /
// sedimentation / final update); the device runs them as three kernels because the scheme's code size (20 k instructions) thrashes the
// instruction cache.  Between the passes the column's LTRUE flag and the two surface precipitation sums travel in levels 1-3 of the W3
// array (an input that only the first section reads).
template <int PH, bool LSYNC, class A, class Cn>
M25_HD inline void column_ph(A &a, const Cn &c, const int nz, const double DT, double &PRECRT, double &SNOWRT) {
  const double PI = c.PI, QSMALL = c.QSMALL, TMELT = c.TMELT, RHOW = c.RHOW, RHOSU = c.RHOSU;
  const double BS = c.BS, BR = c.BR, BG = c.BG, BI = c.BI, BC = c.BC, DS = c.DS, DG = c.DG, DI = c.DI, DCS = c.DCS;
  const double F1S = c.F1S, F2S = c.F2S, F1R = c.F1R, F2R = c.F2R, CPW = c.CPW, AIMM = c.AIMM, MMULT = c.MMULT, RV = c.RV;
  const double *CONS = c.CONS;
  const double XXLV = c.XXLV, XXLS = c.XXLS, CPM = c.CP, EP_2 = c.EP_2;
  const double XLF = XXLS - XXLV;
  const double third = (double)(1.f / 3.f);
  auto GAM = [&](double x) { return gamma_f(x, c.PI, c.SQRTPI); };
  int LTRUE = 0;
  double EPSR = 0., EPSI = 0., EPSS = 0., EPSG = 0., DQSDT = 0.;   // SAVE-less locals of the reference keep their last value across levels
  (void)EPSI; (void)EPSS; (void)EPSG; (void)DQSDT; (void)BC; (void)CPW; (void)AIMM; (void)MMULT; (void)XLF; (void)DCS; (void)F1S; (void)F2S; (void)F1R; (void)F2R;

  if (PH == 0 || PH == 1) {
  for (int K = 1; K <= nz; K++) {
    if (LSYNC) { M25_LEVEL_SYNC(); }   // only when the launch runs whole blocks
    double QC3D = a(QC, K), QI3D = a(QI, K), QNI3D = a(QNI, K), QR3D = a(QR, K), NC3D = a(NC, K), NI3D = a(NI, K), NS3D = a(NS, K),
           NR3D = a(NR, K), QG3D = a(QG, K), NG3D = a(NG, K), T3D = a(T3, K), QV3D = a(QV, K);
    const double PRESK = a(PRES, K), RHOK = a(RHO, K);
    double QC3DTEN = 0., QI3DTEN = 0., QNI3DTEN = 0., QR3DTEN = 0., NC3DTEN = 0., NI3DTEN = 0., NS3DTEN = 0., NR3DTEN = 0., QG3DTEN = 0., NG3DTEN = 0.,
           T3DTEN = 0., QV3DTEN = 0.;
    double aut1d = 0., acc1d = 0., evpc1d = 0., evpr1d = 0., mlt1d = 0., sub1d = 0., dep1d = 0., con1d = 0.;
    // process rates (zeroed for every level at module_mp_graupel.F90:1498-1521, 2014-2037, 3078-3137)
    double PRC = 0., NPRC = 0., NPRC1 = 0., PRA = 0., NPRA = 0., NRAGG = 0., PSMLT = 0., NSMLTS = 0., NSMLTR = 0., EVPMS = 0., PCC = 0., PRE = 0.,
           NSUBC = 0., NSUBR = 0., PRACG = 0., NPRACG = 0., PGMLT = 0., EVPMG = 0., PRACS = 0., NPRACS = 0., NGMLTG = 0., NGMLTR = 0.;
    (void)NSUBC;
    double EVS = mn_(F4(0.99) * PRESK, polysvp(T3D, 0));
    double EIS = mn_(F4(0.99) * PRESK, polysvp(T3D, 1));
    if (EIS > EVS) EIS = EVS;
    const double QVS = EP_2 * EVS / (PRESK - EVS);
    const double QVI = EP_2 * EIS / (PRESK - EIS);
    const double QVQVS = QV3D / QVS;
    const double QVQVSI = QV3D / QVI;
    if (QVQVS < F4(0.9)) {
      if (QR3D < F4(1.E-8)) { QV3D = QV3D + QR3D; T3D = T3D - QR3D * XXLV / CPM; QR3D = 0.; }
      if (QC3D < F4(1.E-8)) { QV3D = QV3D + QC3D; T3D = T3D - QC3D * XXLV / CPM; QC3D = 0.; }
    }
    if (QVQVSI < F4(0.9)) {
      if (QI3D < F4(1.E-8)) { QV3D = QV3D + QI3D; T3D = T3D - QI3D * XXLS / CPM; QI3D = 0.; }
      if (QNI3D < F4(1.E-8)) { QV3D = QV3D + QNI3D; T3D = T3D - QNI3D * XXLS / CPM; QNI3D = 0.; }
      if (QG3D < F4(1.E-8)) { QV3D = QV3D + QG3D; T3D = T3D - QG3D * XXLS / CPM; QG3D = 0.; }
    }
    if (QC3D < QSMALL) { QV3D = QV3D + QC3D; T3D = T3D - QC3D * XXLV / CPM; QC3D = 0.; NC3D = 0.; }
    if (QR3D < QSMALL) { QV3D = QV3D + QR3D; T3D = T3D - QR3D * XXLV / CPM; QR3D = 0.; NR3D = 0.; }
    if (QI3D < QSMALL) { QV3D = QV3D + QI3D; T3D = T3D - QI3D * XXLS / CPM; QI3D = 0.; NI3D = 0.; }
    if (QNI3D < QSMALL) { QV3D = QV3D + QNI3D; T3D = T3D - QNI3D * XXLS / CPM; QNI3D = 0.; NS3D = 0.; }
    if (QG3D < QSMALL) { QV3D = QV3D + QG3D; T3D = T3D - QG3D * XXLS / CPM; QG3D = 0.; NG3D = 0.; }
    a(SQR, K) = 0.; a(SQI, K) = 0.; a(SQNI, K) = 0.; a(SQC, K) = 0.; a(SQG, K) = 0.;
    double LAMI = 0., LAMR = 0., LAMS = 0., LAMG = 0., LAMC = 0., N0I = 0., N0S = 0., N0RR = 0., N0G = 0., PGAM = 0., CDIST1 = 0.;
    bool skip200 = false;
    if (QC3D < QSMALL && QI3D < QSMALL && QNI3D < QSMALL && QR3D < QSMALL && QG3D < QSMALL) {
      if (T3D < TMELT && QVQVSI < F4(0.999)) skip200 = true;
      if (T3D >= TMELT && QVQVS < F4(0.999)) skip200 = true;
    }
    if (!skip200) {
      // Viscosity and the density-corrected fall-speed coefficients (module_mp_graupel.F90:1635-1668; the reference evaluates them for
      // every level before the test above).  A level that takes the GOTO 200 has no hydrometeors and zero tendencies, so the
      // sedimentation never reads its coefficients (they are only used under DUM* >= QSMALL): three x**y per clear-air level saved,
      // results unchanged.
      const double MU = F4(1.496E-6) * M25_POW(T3D, F4(1.5)) / (T3D + F4(120.));
      double DUM = M25_POW(RHOSU / RHOK, F4(0.54));
      const double AIN = M25_POW(RHOSU / RHOK, F4(0.35)) * c.AI;
      const double ARN = DUM * c.AR, ASN = DUM * c.AS;
      const double ACN = c.G * RHOW / (F4(18.) * MU);
      const double AGN = DUM * c.AG;
      a(AIN_, K) = AIN; a(ARN_, K) = ARN; a(ASN_, K) = ASN; a(ACN_, K) = ACN; a(AGN_, K) = AGN;
      const double KAP = F4(1.414E3) * MU;
      const double DV = F4(8.794E-5) * M25_POW(T3D, F4(1.81)) / PRESK;
      const double SC = MU / (RHOK * DV);
      DUM = (RV * (T3D * T3D));
      DQSDT = XXLV * QVS / DUM;
      const double DQSIDT = XXLS * QVI / DUM;
      const double ABI = F4(1.) + DQSIDT * XXLS / CPM;
      const double AB = F4(1.) + DQSDT * XXLV / CPM;
      // size-distribution parameters shared by both temperature branches
      auto rain_dist = [&]() {
        if (QR3D >= QSMALL) {
          LAMR = M25_POW(PI * RHOW * NR3D / QR3D, third);
          N0RR = NR3D * LAMR;
          if (LAMR < c.LAMMINR) { LAMR = c.LAMMINR; N0RR = p4_(LAMR) * QR3D / (PI * RHOW); NR3D = N0RR / LAMR; }
          else if (LAMR > c.LAMMAXR) { LAMR = c.LAMMAXR; N0RR = p4_(LAMR) * QR3D / (PI * RHOW); NR3D = N0RR / LAMR; }
        }
      };
      auto cloud_dist = [&]() {
        if (QC3D >= QSMALL) {
          PGAM = F4(0.0005714) * (NC3D / F4(1.E6) * RHOK) + F4(0.2714);
          PGAM = F4(1.) / (PGAM * PGAM) - F4(1.);
          PGAM = mx_(PGAM, 2.);
          PGAM = mn_(PGAM, 10.);
          LAMC = M25_POW(CONS[26] * NC3D * GAM(PGAM + F4(4.)) / (QC3D * GAM(PGAM + F4(1.))), third);
          const double LAMMIN = (PGAM + F4(1.)) / F4(60.E-6), LAMMAX = (PGAM + F4(1.)) / F4(1.E-6);
          if (LAMC < LAMMIN) {
            LAMC = LAMMIN;
            NC3D = M25_EXP(F4(3.) * M25_LOG(LAMC) + M25_LOG(QC3D) + M25_LOG(GAM(PGAM + F4(1.))) - M25_LOG(GAM(PGAM + F4(4.)))) / CONS[26];
          } else if (LAMC > LAMMAX) {
            LAMC = LAMMAX;
            NC3D = M25_EXP(F4(3.) * M25_LOG(LAMC) + M25_LOG(QC3D) + M25_LOG(GAM(PGAM + F4(1.))) - M25_LOG(GAM(PGAM + F4(4.)))) / CONS[26];
          }
        }
      };
      auto snow_dist = [&]() {
        if (QNI3D >= QSMALL) {
          LAMS = M25_POW(CONS[1] * NS3D / QNI3D, F4(1.) / DS);
          N0S = NS3D * LAMS;
          if (LAMS < c.LAMMINS) { LAMS = c.LAMMINS; N0S = M25_POW(LAMS, DS + F4(1.)) * QNI3D / CONS[1]; NS3D = N0S / LAMS; }
          else if (LAMS > c.LAMMAXS) { LAMS = c.LAMMAXS; N0S = M25_POW(LAMS, DS + F4(1.)) * QNI3D / CONS[1]; NS3D = N0S / LAMS; }
        }
      };
      auto graupel_dist = [&]() {
        if (QG3D >= QSMALL) {
          LAMG = M25_POW(CONS[2] * NG3D / QG3D, F4(1.) / DG);
          N0G = NG3D * LAMG;
          if (LAMG < c.LAMMING) { LAMG = c.LAMMING; N0G = M25_POW(LAMG, DG + F4(1.)) * QG3D / CONS[2]; NG3D = N0G / LAMG; }
          else if (LAMG > c.LAMMAXG) { LAMG = c.LAMMAXG; N0G = M25_POW(LAMG, DG + F4(1.)) * QG3D / CONS[2]; NG3D = N0G / LAMG; }
        }
      };
      // rain-snow / rain-graupel fall-speed pairs (3337-3348, 3395-3406)
      double UMS, UMR, UNS, UNR, UMG, UNG;
      auto epsr_of = [&]() {
        if (QR3D >= QSMALL)
          EPSR = F4(2.) * PI * N0RR * RHOK * DV * (F1R / (LAMR * LAMR) + F2R * M25_POW(ARN * RHOK / MU, F4(0.5)) * M25_POW(SC, third) * CONS[9] / M25_POW(LAMR, CONS[34]));
        else EPSR = 0.;
      };
      auto epss_val = [&]() {
        return F4(2.) * PI * N0S * RHOK * DV * (F1S / (LAMS * LAMS) + F2S * M25_POW(ASN * RHOK / MU, F4(0.5)) * M25_POW(SC, third) * CONS[10] / M25_POW(LAMS, CONS[35]));
      };
      auto epsg_val = [&]() {
        return F4(2.) * PI * N0G * RHOK * DV * (F1S / (LAMG * LAMG) + F2S * M25_POW(AGN * RHOK / MU, F4(0.5)) * M25_POW(SC, third) * CONS[11] / M25_POW(LAMG, CONS[36]));
      };
      bool do_act = true;   // the activation block follows both branches (2517-2845, 4287-4636)
      double wmin_act;
      if (T3D >= TMELT) {
        // ------------------------------------------------------------- warm branch (1815-2510)
        wmin_act = F4(0.20);
        if (QNI3D < F4(1.E-6)) { QR3D = QR3D + QNI3D; NR3D = NR3D + NS3D; T3D = T3D - QNI3D * XLF / CPM; QNI3D = 0.; NS3D = 0.; }
        if (QG3D < F4(1.E-6)) { QR3D = QR3D + QG3D; NR3D = NR3D + NG3D; T3D = T3D - QG3D * XLF / CPM; QG3D = 0.; NG3D = 0.; }
        const bool skip300 = (QC3D < QSMALL && QNI3D < F4(1.E-8) && QR3D < QSMALL && QG3D < F4(1.E-8));
        if (!skip300) {
          NS3D = mx_(0., NS3D); NC3D = mx_(0., NC3D); NR3D = mx_(0., NR3D); NG3D = mx_(0., NG3D);
          rain_dist(); cloud_dist(); snow_dist(); graupel_dist();
          if (QC3D >= F4(1.E-6)) {   // autoconversion, KK2000 (IRAIN = 0)
            PRC = F4(1350.) * M25_POW(QC3D, F4(2.47)) * M25_POW(NC3D / F4(1.e6) * RHOK, F4(-1.79));
            NPRC1 = PRC / CONS[29];
            NPRC = PRC / (QC3D / NC3D);
            NPRC = mn_(NPRC, NC3D / DT);
          }
          if (QR3D >= F4(1.E-8) && QNI3D >= F4(1.E-8)) {
            UMS = ASN * CONS[3] / M25_POW(LAMS, BS); UMR = ARN * CONS[4] / M25_POW(LAMR, BR);
            UNS = ASN * CONS[5] / M25_POW(LAMS, BS); UNR = ARN * CONS[6] / M25_POW(LAMR, BR);
            DUM = M25_POW(RHOSU / RHOK, F4(0.54));
            UMS = mn_(UMS, F4(1.2) * DUM); UNS = mn_(UNS, F4(1.2) * DUM); UMR = mn_(UMR, F4(9.1) * DUM); UNR = mn_(UNR, F4(9.1) * DUM);
            PRACS = CONS[31] * (M25_POW(p2_(F4(1.2) * UMR - F4(0.95) * UMS) + F4(0.08) * UMS * UMR, F4(0.5)) * RHOK * N0RR * N0S / p3_(LAMS) *
                                (F4(5.) / (p3_(LAMS) * LAMR) + F4(2.) / (p2_(LAMS) * p2_(LAMR)) + F4(0.5) / (LAMS * p3_(LAMR))));
          }
          if (QR3D >= F4(1.E-8) && QG3D >= F4(1.E-8)) {
            UMG = AGN * CONS[7] / M25_POW(LAMG, BG); UMR = ARN * CONS[4] / M25_POW(LAMR, BR);
            UNG = AGN * CONS[8] / M25_POW(LAMG, BG); UNR = ARN * CONS[6] / M25_POW(LAMR, BR);
            DUM = M25_POW(RHOSU / RHOK, F4(0.54));
            UMG = mn_(UMG, F4(20.) * DUM); UNG = mn_(UNG, F4(20.) * DUM); UMR = mn_(UMR, F4(9.1) * DUM); UNR = mn_(UNR, F4(9.1) * DUM);
            PRACG = CONS[41] * (M25_POW(p2_(F4(1.2) * UMR - F4(0.95) * UMG) + F4(0.08) * UMG * UMR, F4(0.5)) * RHOK * N0RR * N0G / p3_(LAMR) *
                                (F4(5.) / (p3_(LAMR) * LAMG) + F4(2.) / (p2_(LAMR) * p2_(LAMG)) + F4(0.5) / (LAMR * p3_(LAMG))));
            DUM = PRACG / F4(5.2E-7);
            NPRACG = CONS[32] * RHOK * M25_POW(F4(1.7) * p2_(UNR - UNG) + F4(0.3) * UNR * UNG, F4(0.5)) * N0RR * N0G *
                     (F4(1.) / (p3_(LAMR) * LAMG) + F4(1.) / (p2_(LAMR) * p2_(LAMG)) + F4(1.) / (LAMR * p3_(LAMG)));
            NPRACG = mx_(NPRACG - DUM, 0.);
          }
          if (QR3D >= F4(1.E-8) && QC3D >= F4(1.E-8)) {
            DUM = (QC3D * QR3D);
            PRA = F4(67.) * M25_POW(DUM, F4(1.15));
            NPRA = PRA / (QC3D / NC3D);
          }
          if (QR3D >= F4(1.E-8)) {
            const double dum1 = F4(300.e-6);
            if (F4(1.) / LAMR < dum1) DUM = 1.;
            else DUM = F4(2.) - M25_EXP(F4(2300.) * (F4(1.) / LAMR - dum1));
            NRAGG = F4(-5.78) * DUM * NR3D * QR3D * RHOK;
          }
          epsr_of();
          if (QV3D < QVS) { PRE = EPSR * (QV3D - QVS) / AB; PRE = mn_(PRE, 0.); } else PRE = 0.;
          if (QNI3D >= F4(1.E-8)) {
            DUM = -CPW / XLF * mx_(T3D - TMELT, 0.) * PRACS;
            PSMLT = F4(2.) * PI * N0S * KAP * mn_(TMELT - T3D, 0.) / XLF * RHOK *
                        (F1S / (LAMS * LAMS) + F2S * M25_POW(ASN * RHOK / MU, F4(0.5)) * M25_POW(SC, third) * CONS[10] / M25_POW(LAMS, CONS[35])) + DUM;
            if (QVQVS < F4(1.)) {
              EPSS = epss_val();
              EVPMS = (QV3D - QVS) * EPSS / AB;
              EVPMS = mx_(EVPMS, PSMLT);
              PSMLT = PSMLT - EVPMS;
            }
          }
          if (QG3D >= F4(1.E-8)) {
            DUM = -CPW / XLF * mx_(T3D - TMELT, 0.) * PRACG;
            PGMLT = F4(2.) * PI * N0G * KAP * (TMELT - T3D) / XLF * RHOK *
                        (F1S / (LAMG * LAMG) + F2S * M25_POW(AGN * RHOK / MU, F4(0.5)) * M25_POW(SC, third) * CONS[11] / M25_POW(LAMG, CONS[36])) + DUM;
            if (QVQVS < F4(1.)) {
              EPSG = epsg_val();
              EVPMG = (QV3D - QVS) * EPSG / AB;
              EVPMG = mx_(EVPMG, PGMLT);
              PGMLT = PGMLT - EVPMG;
            }
          }
          PRACG = 0.; PRACS = 0.;
          // conservation of water (2347-2399)
          DUM = (PRC + PRA) * DT;
          if (DUM > QC3D && QC3D >= QSMALL) { const double RATIO = QC3D / DUM; PRC = PRC * RATIO; PRA = PRA * RATIO; }
          DUM = (-PSMLT - EVPMS + PRACS) * DT;
          if (DUM > QNI3D && QNI3D >= QSMALL) { const double RATIO = QNI3D / DUM; PSMLT = PSMLT * RATIO; EVPMS = EVPMS * RATIO; PRACS = PRACS * RATIO; }
          DUM = (-PGMLT - EVPMG + PRACG) * DT;
          if (DUM > QG3D && QG3D >= QSMALL) { const double RATIO = QG3D / DUM; PGMLT = PGMLT * RATIO; EVPMG = EVPMG * RATIO; PRACG = PRACG * RATIO; }
          DUM = (-PRACS - PRACG - PRE - PRA - PRC + PSMLT + PGMLT) * DT;
          if (DUM > QR3D && QR3D >= QSMALL) {
            const double RATIO = (QR3D / DT + PRACS + PRACG + PRA + PRC - PSMLT - PGMLT) / (-PRE);
            PRE = PRE * RATIO;
          }
          QV3DTEN = QV3DTEN + (-PRE - EVPMS - EVPMG);
          T3DTEN = T3DTEN + (PRE * XXLV + (EVPMS + EVPMG) * XXLS + (PSMLT + PGMLT - PRACS - PRACG) * XLF) / CPM;
          QC3DTEN = QC3DTEN + (-PRA - PRC);
          QR3DTEN = QR3DTEN + (PRE + PRA + PRC - PSMLT - PGMLT + PRACS + PRACG);
          QNI3DTEN = QNI3DTEN + (PSMLT + EVPMS - PRACS);
          QG3DTEN = QG3DTEN + (PGMLT + EVPMG - PRACG);
          NC3DTEN = NC3DTEN + (-NPRA - NPRC);
          NR3DTEN = NR3DTEN + (NPRC1 + NRAGG - NPRACG);
          if (PRE < 0.) { DUM = PRE * DT / QR3D; DUM = mx_(-1., DUM); NSUBR = DUM * NR3D / DT; }
          if (EVPMS + PSMLT < 0.) { DUM = (EVPMS + PSMLT) * DT / QNI3D; DUM = mx_(-1., DUM); NSMLTS = DUM * NS3D / DT; }
          if (PSMLT < 0.) { DUM = PSMLT * DT / QNI3D; DUM = mx_(-1., DUM); NSMLTR = DUM * NS3D / DT; }
          if (EVPMG + PGMLT < 0.) { DUM = (EVPMG + PGMLT) * DT / QG3D; DUM = mx_(-1., DUM); NGMLTG = DUM * NG3D / DT; }
          if (PGMLT < 0.) { DUM = PGMLT * DT / QG3D; DUM = mx_(-1., DUM); NGMLTR = DUM * NG3D / DT; }
          NS3DTEN = NS3DTEN + (NSMLTS);
          NG3DTEN = NG3DTEN + (NGMLTG);
          NR3DTEN = NR3DTEN + (NSUBR - NSMLTR - NGMLTR);
        }
        // 300 CONTINUE; ISATADJ = 1: no condensation here
        aut1d = PRC; acc1d = PRA;
        mlt1d = -PSMLT - PGMLT + PRACS + PRACG;
        evpr1d = -PRE - EVPMS - EVPMG;
        if (PCC < 0.) evpc1d = -PCC; else if (PCC > 0.) con1d = PCC;
      } else {
        // ------------------------------------------------------------- cold branch (2864-4281)
        wmin_act = F4(0.10);
        double MNUCCC = 0., NNUCCC = 0., NSAGG = 0., PSACWS = 0., NPSACWS = 0., PSACWI = 0., NPSACWI = 0., NMULTS = 0., QMULTS = 0., NMULTR = 0., QMULTR = 0.,
               NMULTG = 0., QMULTG = 0., NMULTRG = 0., QMULTRG = 0., MNUCCR = 0., NNUCCR = 0., PRCI = 0., NPRCI = 0., PRAI = 0., NPRAI = 0., NNUCCD = 0., MNUCCD = 0.,
               PRD = 0., PRDS = 0., EPRD = 0., EPRDS = 0., NSUBI = 0., NSUBS = 0., PIACR = 0., NIACR = 0., PRACI = 0., PIACRS = 0., NIACRS = 0., PRACIS = 0.,
               PSACR = 0., PSACWG = 0., PGSACW = 0., PGRACS = 0., PRDG = 0., EPRDG = 0., NPSACWG = 0., NSCNG = 0., NGRACS = 0., NSUBG = 0.;
        NI3D = mx_(0., NI3D); NS3D = mx_(0., NS3D); NC3D = mx_(0., NC3D); NR3D = mx_(0., NR3D); NG3D = mx_(0., NG3D);
        if (QI3D >= QSMALL) {
          LAMI = M25_POW(CONS[12] * NI3D / QI3D, F4(1.) / DI);
          N0I = NI3D * LAMI;
          if (LAMI < c.LAMMINI) { LAMI = c.LAMMINI; N0I = M25_POW(LAMI, DI + F4(1.)) * QI3D / CONS[12]; NI3D = N0I / LAMI; }
          else if (LAMI > c.LAMMAXI) { LAMI = c.LAMMAXI; N0I = M25_POW(LAMI, DI + F4(1.)) * QI3D / CONS[12]; NI3D = N0I / LAMI; }
        }
        rain_dist(); cloud_dist();
        if (QC3D >= QSMALL) CDIST1 = NC3D / GAM(PGAM + F4(1.));
        snow_dist(); graupel_dist();
        if (QC3D >= QSMALL && T3D < F4(269.15)) {   // contact + immersion freezing of droplets (3145-3192)
          const double NACNT = M25_EXP(F4(-2.80) + F4(0.262) * (TMELT - T3D)) * F4(1000.);
          DUM = F4(7.37) * T3D / (F4(288.) * F4(10.) * PRESK) / F4(100.);
          const double DAP = CONS[37] * T3D * (F4(1.) + DUM / c.RIN) / MU;
          MNUCCC = CONS[38] * DAP * NACNT * M25_EXP(M25_LOG(CDIST1) + M25_LOG(GAM(PGAM + F4(5.))) - F4(4.) * M25_LOG(LAMC));
          NNUCCC = F4(2.) * PI * DAP * NACNT * CDIST1 * GAM(PGAM + F4(2.)) / LAMC;
          MNUCCC = MNUCCC + CONS[39] * M25_EXP(M25_LOG(CDIST1) + M25_LOG(GAM(F4(7.) + PGAM)) - F4(6.) * M25_LOG(LAMC)) * M25_EXP(AIMM * (TMELT - T3D));
          NNUCCC = NNUCCC + CONS[40] * M25_EXP(M25_LOG(CDIST1) + M25_LOG(GAM(PGAM + F4(4.))) - F4(3.) * M25_LOG(LAMC)) * M25_EXP(AIMM * (TMELT - T3D));
          NNUCCC = mn_(NNUCCC, NC3D / DT);
        }
        if (QC3D >= F4(1.E-6)) {
          PRC = F4(1350.) * M25_POW(QC3D, F4(2.47)) * M25_POW(NC3D / F4(1.e6) * RHOK, F4(-1.79));
          NPRC1 = PRC / CONS[29];
          NPRC = PRC / (QC3D / NC3D);
          NPRC = mn_(NPRC, NC3D / DT);
        }
        if (QNI3D >= F4(1.E-8))
          NSAGG = CONS[15] * ASN * M25_POW(RHOK, (F4(2.) + BS) / F4(3.)) * M25_POW(QNI3D, (F4(2.) + BS) / F4(3.)) * M25_POW(NS3D * RHOK, (F4(4.) - BS) / F4(3.)) / (RHOK);
        if (QNI3D >= F4(1.E-8) && QC3D >= QSMALL) {
          PSACWS = CONS[13] * ASN * QC3D * RHOK * N0S / M25_POW(LAMS, BS + F4(3.));
          NPSACWS = CONS[13] * ASN * NC3D * RHOK * N0S / M25_POW(LAMS, BS + F4(3.));
        }
        if (QG3D >= F4(1.E-8) && QC3D >= QSMALL) {
          PSACWG = CONS[14] * AGN * QC3D * RHOK * N0G / M25_POW(LAMG, BG + F4(3.));
          NPSACWG = CONS[14] * AGN * NC3D * RHOK * N0G / M25_POW(LAMG, BG + F4(3.));
        }
        if (QI3D >= F4(1.E-8) && QC3D >= QSMALL) {
          if (F4(1.) / LAMI >= F4(100.E-6)) {
            PSACWI = CONS[16] * AIN * QC3D * RHOK * N0I / M25_POW(LAMI, BI + F4(3.));
            NPSACWI = CONS[16] * AIN * NC3D * RHOK * N0I / M25_POW(LAMI, BI + F4(3.));
          }
        }
        if (QR3D >= F4(1.E-8) && QNI3D >= F4(1.E-8)) {
          UMS = ASN * CONS[3] / M25_POW(LAMS, BS); UMR = ARN * CONS[4] / M25_POW(LAMR, BR);
          UNS = ASN * CONS[5] / M25_POW(LAMS, BS); UNR = ARN * CONS[6] / M25_POW(LAMR, BR);
          DUM = M25_POW(RHOSU / RHOK, F4(0.54));
          UMS = mn_(UMS, F4(1.2) * DUM); UNS = mn_(UNS, F4(1.2) * DUM); UMR = mn_(UMR, F4(9.1) * DUM); UNR = mn_(UNR, F4(9.1) * DUM);
          PRACS = CONS[41] * (M25_POW(p2_(F4(1.2) * UMR - F4(0.95) * UMS) + F4(0.08) * UMS * UMR, F4(0.5)) * RHOK * N0RR * N0S / p3_(LAMR) *
                              (F4(5.) / (p3_(LAMR) * LAMS) + F4(2.) / (p2_(LAMR) * p2_(LAMS)) + F4(0.5) / (LAMR * p3_(LAMS))));
          NPRACS = CONS[32] * RHOK * M25_POW(F4(1.7) * p2_(UNR - UNS) + F4(0.3) * UNR * UNS, F4(0.5)) * N0RR * N0S *
                   (F4(1.) / (p3_(LAMR) * LAMS) + F4(1.) / (p2_(LAMR) * p2_(LAMS)) + F4(1.) / (LAMR * p3_(LAMS)));
          PRACS = mn_(PRACS, QR3D / DT);
          if (QNI3D >= F4(0.1E-3) && QR3D >= F4(0.1E-3))
            PSACR = CONS[31] * (M25_POW(p2_(F4(1.2) * UMR - F4(0.95) * UMS) + F4(0.08) * UMS * UMR, F4(0.5)) * RHOK * N0RR * N0S / p3_(LAMS) *
                                (F4(5.) / (p3_(LAMS) * LAMR) + F4(2.) / (p2_(LAMS) * p2_(LAMR)) + F4(0.5) / (LAMS * p3_(LAMR))));
        }
        if (QR3D >= F4(1.E-8) && QG3D >= F4(1.E-8)) {
          UMG = AGN * CONS[7] / M25_POW(LAMG, BG); UMR = ARN * CONS[4] / M25_POW(LAMR, BR);
          UNG = AGN * CONS[8] / M25_POW(LAMG, BG); UNR = ARN * CONS[6] / M25_POW(LAMR, BR);
          DUM = M25_POW(RHOSU / RHOK, F4(0.54));
          UMG = mn_(UMG, F4(20.) * DUM); UNG = mn_(UNG, F4(20.) * DUM); UMR = mn_(UMR, F4(9.1) * DUM); UNR = mn_(UNR, F4(9.1) * DUM);
          PRACG = CONS[41] * (M25_POW(p2_(F4(1.2) * UMR - F4(0.95) * UMG) + F4(0.08) * UMG * UMR, F4(0.5)) * RHOK * N0RR * N0G / p3_(LAMR) *
                              (F4(5.) / (p3_(LAMR) * LAMG) + F4(2.) / (p2_(LAMR) * p2_(LAMG)) + F4(0.5) / (LAMR * p3_(LAMG))));
          NPRACG = CONS[32] * RHOK * M25_POW(F4(1.7) * p2_(UNR - UNG) + F4(0.3) * UNR * UNG, F4(0.5)) * N0RR * N0G *
                   (F4(1.) / (p3_(LAMR) * LAMG) + F4(1.) / (p2_(LAMR) * p2_(LAMG)) + F4(1.) / (LAMR * p3_(LAMG)));
          PRACG = mn_(PRACG, QR3D / DT);
        }
        // rime splintering, Hallet-Mossop (3441-3555)
        auto fmult_of = [&]() {
          double FMULT = 0.;
          if (T3D > F4(270.16)) FMULT = 0.;
          else if (T3D <= F4(270.16) && T3D > F4(268.16)) FMULT = (F4(270.16) - T3D) / F4(2.);
          else if (T3D >= F4(265.16) && T3D <= F4(268.16)) FMULT = (T3D - F4(265.16)) / F4(3.);
          else if (T3D < F4(265.16)) FMULT = 0.;
          return FMULT;
        };
        if (QNI3D >= F4(0.1E-3)) {
          if (QC3D >= F4(0.5E-3) || QR3D >= F4(0.1E-3)) {
            if (PSACWS > 0. || PRACS > 0.) {
              if (T3D < F4(270.16) && T3D > F4(265.16)) {
                const double FMULT = fmult_of();
                if (PSACWS > 0.) { NMULTS = F4(35.E4) * PSACWS * FMULT * F4(1000.); QMULTS = NMULTS * MMULT; QMULTS = mn_(QMULTS, PSACWS); PSACWS = PSACWS - QMULTS; }
                if (PRACS > 0.) { NMULTR = F4(35.E4) * PRACS * FMULT * F4(1000.); QMULTR = NMULTR * MMULT; QMULTR = mn_(QMULTR, PRACS); PRACS = PRACS - QMULTR; }
              }
            }
          }
        }
        if (QG3D >= F4(0.1E-3)) {
          if (QC3D >= F4(0.5E-3) || QR3D >= F4(0.1E-3)) {
            if (PSACWG > 0. || PRACG > 0.) {
              if (T3D < F4(270.16) && T3D > F4(265.16)) {
                const double FMULT = fmult_of();
                if (PSACWG > 0.) { NMULTG = F4(35.E4) * PSACWG * FMULT * F4(1000.); QMULTG = NMULTG * MMULT; QMULTG = mn_(QMULTG, PSACWG); PSACWG = PSACWG - QMULTG; }
                if (PRACG > 0.) { NMULTRG = F4(35.E4) * PRACG * FMULT * F4(1000.); QMULTRG = NMULTRG * MMULT; QMULTRG = mn_(QMULTRG, PRACG); PRACG = PRACG - QMULTRG; }
              }
            }
          }
        }
        // conversion of rimed snow to graupel (3567-3612)
        if (PSACWS > 0.) {
          if (QNI3D >= F4(0.1E-3) && QC3D >= F4(0.5E-3)) {
            PGSACW = mn_(PSACWS, CONS[17] * DT * N0S * QC3D * QC3D * ASN * ASN / (RHOK * M25_POW(LAMS, F4(2.) * BS + F4(2.))));
            DUM = mx_(c.RHOSN / (c.RHOG - c.RHOSN) * PGSACW, 0.);
            NSCNG = DUM / c.MG0 * RHOK;
            NSCNG = mn_(NSCNG, NS3D / DT);
            PSACWS = PSACWS - PGSACW;
          }
        }
        if (PRACS > 0.) {
          if (QNI3D >= F4(0.1E-3) && QR3D >= F4(0.1E-3)) {
            DUM = CONS[18] * p3_(F4(4.) / LAMS) * p3_(F4(4.) / LAMS) /
                  (CONS[18] * p3_(F4(4.) / LAMS) * p3_(F4(4.) / LAMS) + CONS[19] * p3_(F4(4.) / LAMR) * p3_(F4(4.) / LAMR));
            DUM = mn_(DUM, 1.); DUM = mx_(DUM, 0.);
            PGRACS = (F4(1.) - DUM) * PRACS;
            NGRACS = (F4(1.) - DUM) * NPRACS;
            NGRACS = mn_(NGRACS, NR3D / DT); NGRACS = mn_(NGRACS, NS3D / DT);
            PRACS = PRACS - PGRACS;
            NPRACS = NPRACS - NGRACS;
            PSACR = PSACR * (F4(1.) - DUM);
          }
        }
        if (T3D < F4(269.15) && QR3D >= QSMALL) {   // freezing of rain (3619-3630)
          MNUCCR = CONS[20] * NR3D * M25_EXP(AIMM * (TMELT - T3D)) / p3_(LAMR) / p3_(LAMR);
          NNUCCR = PI * NR3D * c.BIMM * M25_EXP(AIMM * (TMELT - T3D)) / p3_(LAMR);
          NNUCCR = mn_(NNUCCR, NR3D / DT);
        }
        if (QR3D >= F4(1.E-8) && QC3D >= F4(1.E-8)) {
          DUM = (QC3D * QR3D);
          PRA = F4(67.) * M25_POW(DUM, F4(1.15));
          NPRA = PRA / (QC3D / NC3D);
        }
        if (QR3D >= F4(1.E-8)) {
          const double dum1 = F4(300.e-6);
          if (F4(1.) / LAMR < dum1) DUM = 1.;
          else DUM = F4(2.) - M25_EXP(F4(2300.) * (F4(1.) / LAMR - dum1));
          NRAGG = F4(-5.78) * DUM * NR3D * QR3D * RHOK;
        }
        if (QI3D >= F4(1.E-8) && QVQVSI >= F4(1.)) {   // autoconversion of cloud ice to snow (3693-3703)
          NPRCI = CONS[21] * (QV3D - QVI) * RHOK * N0I * M25_EXP(-LAMI * DCS) * DV / ABI;
          PRCI = CONS[22] * NPRCI;
          NPRCI = mn_(NPRCI, NI3D / DT);
        }
        if (QNI3D >= F4(1.E-8) && QI3D >= QSMALL) {
          PRAI = CONS[23] * ASN * QI3D * RHOK * N0S / M25_POW(LAMS, BS + F4(3.));
          NPRAI = CONS[23] * ASN * NI3D * RHOK * N0S / M25_POW(LAMS, BS + F4(3.));
          NPRAI = mn_(NPRAI, NI3D / DT);
        }
        if (QR3D >= F4(1.E-8) && QI3D >= F4(1.E-8) && T3D <= TMELT) {
          if (QR3D >= F4(0.1E-3)) {
            NIACR = CONS[24] * NI3D * N0RR * ARN / M25_POW(LAMR, BR + F4(3.)) * RHOK;
            PIACR = CONS[25] * NI3D * N0RR * ARN / M25_POW(LAMR, BR + F4(3.)) / p3_(LAMR) * RHOK;
            PRACI = CONS[24] * QI3D * N0RR * ARN / M25_POW(LAMR, BR + F4(3.)) * RHOK;
            NIACR = mn_(NIACR, NR3D / DT); NIACR = mn_(NIACR, NI3D / DT);
          } else {
            NIACRS = CONS[24] * NI3D * N0RR * ARN / M25_POW(LAMR, BR + F4(3.)) * RHOK;
            PIACRS = CONS[25] * NI3D * N0RR * ARN / M25_POW(LAMR, BR + F4(3.)) / p3_(LAMR) * RHOK;
            PRACIS = CONS[24] * QI3D * N0RR * ARN / M25_POW(LAMR, BR + F4(3.)) * RHOK;
            NIACRS = mn_(NIACRS, NR3D / DT); NIACRS = mn_(NIACRS, NI3D / DT);
          }
        }
        // ice nucleation, Cooper (INUC = 0, 3782-3804)
        if ((QVQVS >= F4(0.999) && T3D <= F4(265.15)) || QVQVSI >= F4(1.08)) {
          double kc2 = F4(0.005) * M25_EXP(F4(0.304) * (TMELT - T3D)) * F4(1000.);
          kc2 = mn_(kc2, F4(500.e3));
          kc2 = mx_(kc2 / RHOK, 0.);
          if (kc2 > NI3D + NS3D + NG3D) { NNUCCD = (kc2 - NI3D - NS3D - NG3D) / DT; MNUCCD = NNUCCD * c.MI0; }
        }
        // deposition / sublimation / evaporation (3842-3988)
        if (QI3D >= QSMALL) EPSI = F4(2.) * PI * N0I * RHOK * DV / (LAMI * LAMI); else EPSI = 0.;
        if (QNI3D >= QSMALL) EPSS = epss_val(); else EPSS = 0.;
        if (QG3D >= QSMALL) EPSG = epsg_val(); else EPSG = 0.;
        epsr_of();
        if (QI3D >= QSMALL) {
          DUM = (F4(1.) - M25_EXP(-LAMI * DCS) * (F4(1.) + LAMI * DCS));
          PRD = EPSI * (QV3D - QVI) / ABI * DUM;
        } else DUM = 0.;
        if (QNI3D >= QSMALL) PRDS = EPSS * (QV3D - QVI) / ABI + EPSI * (QV3D - QVI) / ABI * (F4(1.) - DUM);
        else PRD = PRD + EPSI * (QV3D - QVI) / ABI * (F4(1.) - DUM);
        PRDG = EPSG * (QV3D - QVI) / ABI;
        if (QV3D < QVS) { PRE = EPSR * (QV3D - QVS) / AB; PRE = mn_(PRE, 0.); } else PRE = 0.;
        DUM = (QV3D - QVI) / DT;
        const double FUDGEF = F4(0.9999);
        const double SUM_DEP = PRD + PRDS + MNUCCD + PRDG;
        if ((DUM > 0. && SUM_DEP > DUM * FUDGEF) || (DUM < 0. && SUM_DEP < DUM * FUDGEF)) {
          MNUCCD = FUDGEF * MNUCCD * DUM / SUM_DEP;
          PRD = FUDGEF * PRD * DUM / SUM_DEP;
          PRDS = FUDGEF * PRDS * DUM / SUM_DEP;
          PRDG = FUDGEF * PRDG * DUM / SUM_DEP;
        }
        if (PRD < 0.) { EPRD = PRD; PRD = 0.; }
        if (PRDS < 0.) { EPRDS = PRDS; PRDS = 0.; }
        if (PRDG < 0.) { EPRDG = PRDG; PRDG = 0.; }
        // conservation of water (4044-4144), ILIQ = 0, IGRAUP = 0
        DUM = (PRC + PRA + MNUCCC + PSACWS + PSACWI + QMULTS + PSACWG + PGSACW + QMULTG) * DT;
        if (DUM > QC3D && QC3D >= QSMALL) {
          const double RATIO = QC3D / DUM;
          PRC = PRC * RATIO; PRA = PRA * RATIO; MNUCCC = MNUCCC * RATIO; PSACWS = PSACWS * RATIO; PSACWI = PSACWI * RATIO; QMULTS = QMULTS * RATIO;
          QMULTG = QMULTG * RATIO; PSACWG = PSACWG * RATIO; PGSACW = PGSACW * RATIO;
        }
        DUM = (-PRD - MNUCCC + PRCI + PRAI - QMULTS - QMULTG - QMULTR - QMULTRG - MNUCCD + PRACI + PRACIS - EPRD - PSACWI) * DT;
        if (DUM > QI3D && QI3D >= QSMALL) {
          const double RATIO = (QI3D / DT + PRD + MNUCCC + QMULTS + QMULTG + QMULTR + QMULTRG + MNUCCD + PSACWI) / (PRCI + PRAI + PRACI + PRACIS - EPRD);
          PRCI = PRCI * RATIO; PRAI = PRAI * RATIO; PRACI = PRACI * RATIO; PRACIS = PRACIS * RATIO; EPRD = EPRD * RATIO;
        }
        DUM = ((PRACS - PRE) + (QMULTR + QMULTRG - PRC) + (MNUCCR - PRA) + PIACR + PIACRS + PGRACS + PRACG) * DT;
        if (DUM > QR3D && QR3D >= QSMALL) {
          const double RATIO = (QR3D / DT + PRC + PRA) / (-PRE + QMULTR + QMULTRG + PRACS + MNUCCR + PIACR + PIACRS + PGRACS + PRACG);
          PRE = PRE * RATIO; PRACS = PRACS * RATIO; QMULTR = QMULTR * RATIO; QMULTRG = QMULTRG * RATIO; MNUCCR = MNUCCR * RATIO; PIACR = PIACR * RATIO;
          PIACRS = PIACRS * RATIO; PGRACS = PGRACS * RATIO; PRACG = PRACG * RATIO;
        }
        DUM = (-PRDS - PSACWS - PRAI - PRCI - PRACS - EPRDS + PSACR - PIACRS - PRACIS) * DT;
        if (DUM > QNI3D && QNI3D >= QSMALL) {
          const double RATIO = (QNI3D / DT + PRDS + PSACWS + PRAI + PRCI + PRACS + PIACRS + PRACIS) / (-EPRDS + PSACR);
          EPRDS = EPRDS * RATIO; PSACR = PSACR * RATIO;
        }
        DUM = (-PSACWG - PRACG - PGSACW - PGRACS - PRDG - MNUCCR - EPRDG - PIACR - PRACI - PSACR) * DT;
        if (DUM > QG3D && QG3D >= QSMALL) {
          const double RATIO = (QG3D / DT + PSACWG + PRACG + PGSACW + PGRACS + PRDG + MNUCCR + PSACR + PIACR + PRACI) / (-EPRDG);
          EPRDG = EPRDG * RATIO;
        }
        QV3DTEN = QV3DTEN + (-PRE - PRD - PRDS - MNUCCD - EPRD - EPRDS - PRDG - EPRDG);
        T3DTEN = T3DTEN + (PRE * XXLV + (PRD + PRDS + MNUCCD + EPRD + EPRDS + PRDG + EPRDG) * XXLS +
                           (PSACWS + PSACWI + MNUCCC + MNUCCR + QMULTS + QMULTG + QMULTR + QMULTRG + PRACS + PSACWG + PRACG + PGSACW + PGRACS + PIACR + PIACRS) * XLF) / CPM;
        QC3DTEN = QC3DTEN + (-PRA - PRC - MNUCCC + PCC - PSACWS - PSACWI - QMULTS - QMULTG - PSACWG - PGSACW);
        QI3DTEN = QI3DTEN + (PRD + EPRD + PSACWI + MNUCCC - PRCI - PRAI + QMULTS + QMULTG + QMULTR + QMULTRG + MNUCCD - PRACI - PRACIS);
        QR3DTEN = QR3DTEN + (PRE + PRA + PRC - PRACS - MNUCCR - QMULTR - QMULTRG - PIACR - PIACRS - PRACG - PGRACS);
        QNI3DTEN = QNI3DTEN + (PRAI + PSACWS + PRDS + PRACS + PRCI + EPRDS - PSACR + PIACRS + PRACIS);
        NS3DTEN = NS3DTEN + (NSAGG + NPRCI - NSCNG - NGRACS + NIACRS);
        QG3DTEN = QG3DTEN + (PRACG + PSACWG + PGSACW + PGRACS + PRDG + EPRDG + MNUCCR + PIACR + PRACI + PSACR);
        NG3DTEN = NG3DTEN + (NSCNG + NGRACS + NNUCCR + NIACR);
        NC3DTEN = NC3DTEN + (-NNUCCC - NPSACWS - NPRA - NPRC - NPSACWI - NPSACWG);
        NI3DTEN = NI3DTEN + (NNUCCC - NPRCI - NPRAI + NMULTS + NMULTG + NMULTR + NMULTRG + NNUCCD - NIACR - NIACRS);
        NR3DTEN = NR3DTEN + (NPRC1 - NPRACS - NNUCCR + NRAGG - NIACR - NIACRS - NPRACG - NGRACS);
        if (EPRD < 0.) { DUM = EPRD * DT / QI3D; DUM = mx_(-1., DUM); NSUBI = DUM * NI3D / DT; }
        if (EPRDS < 0.) { DUM = EPRDS * DT / QNI3D; DUM = mx_(-1., DUM); NSUBS = DUM * NS3D / DT; }
        if (PRE < 0.) { DUM = PRE * DT / QR3D; DUM = mx_(-1., DUM); NSUBR = DUM * NR3D / DT; }
        if (EPRDG < 0.) { DUM = EPRDG * DT / QG3D; DUM = mx_(-1., DUM); NSUBG = DUM * NG3D / DT; }
        NI3DTEN = NI3DTEN + NSUBI;
        NS3DTEN = NS3DTEN + NSUBS;
        NG3DTEN = NG3DTEN + NSUBG;
        NR3DTEN = NR3DTEN + NSUBR;
        aut1d = PRC; acc1d = PRA; evpr1d = -PRE;
        if (PCC < 0.) evpc1d = -PCC; else if (PCC > 0.) con1d = PCC;
        sub1d = -EPRD - EPRDS - EPRDG;
        dep1d = PRD + PRDS + MNUCCD + PRDG;
      }
      // ---- droplet activation (INUM = 0, ISUB = 0, IBASE = 2, IACT = 2): 2517-2845 (warm) == 4287-4636 (cold)
      if (do_act && QC3D + QC3DTEN * DT >= QSMALL) {
        DUM = a(W3, K) + a(WVAR, K);
        DUM = mx_(DUM, wmin_act);
        if (DUM >= F4(0.001)) {
          double DUM2 = activate2(c, DUM, T3D, EVS, DV, KAP, PRESK, RHOK);
          DUM2 = (DUM2 - NC3D) / DT;
          DUM2 = mx_(0., DUM2);
          NC3DTEN = NC3DTEN + DUM2;
        }
      }
      LTRUE = 1;
    }
    // 200 CONTINUE: store the level
    a(QC, K) = QC3D; a(QI, K) = QI3D; a(QNI, K) = QNI3D; a(QR, K) = QR3D; a(NC, K) = NC3D; a(NI, K) = NI3D; a(NS, K) = NS3D; a(NR, K) = NR3D;
    a(QG, K) = QG3D; a(NG, K) = NG3D; a(T3, K) = T3D; a(QV, K) = QV3D;
    a(TQC, K) = QC3DTEN; a(TQI, K) = QI3DTEN; a(TQNI, K) = QNI3DTEN; a(TQR, K) = QR3DTEN; a(TNC, K) = NC3DTEN; a(TNI, K) = NI3DTEN; a(TNS, K) = NS3DTEN;
    a(TNR, K) = NR3DTEN; a(TQG, K) = QG3DTEN; a(TNG, K) = NG3DTEN; a(TT, K) = T3DTEN; a(TQV, K) = QV3DTEN;
    a(LAMI_, K) = LAMI;
    a(AUT, K) = aut1d; a(ACC, K) = acc1d; a(EVPC, K) = evpc1d; a(EVPR, K) = evpr1d; a(MLT, K) = mlt1d; a(SUB, K) = sub1d; a(DEP, K) = dep1d; a(CON, K) = con1d;
  }

  if (PH == 1) { a(W3, 1) = (double)LTRUE; return; }
  }   // PH 0 / 1
  if (PH >= 2) LTRUE = (int)a(W3, 1);
  if (PH == 0 || PH == 2) { PRECRT = 0.; SNOWRT = 0.; }
  if (PH == 3) { PRECRT = a(W3, 2); SNOWRT = a(W3, 3); }
  if (LTRUE == 0) {   // GOTO 400
    if (PH == 2) { a(W3, 2) = 0.; a(W3, 3) = 0.; }
    return;
  }
  if (PH == 0 || PH == 2) {

  // ------------------------------------------------------------------- sedimentation (4796-5127)
  int NSTEP = 1;
  double frn = 0., fin = 0., fnin = 0., fsn = 0., fnsn = 0., fnrn = 0., fcn = 0., fncn = 0., fgn = 0., fngn = 0.;   // F*(K+1)
  for (int K = nz; K >= 1; K--) {
    const double RHOK = a(RHO, K);
    double DUMI = a(QI, K) + a(TQI, K) * DT, DUMQS = a(QNI, K) + a(TQNI, K) * DT, DUMR = a(QR, K) + a(TQR, K) * DT;
    double DUMFNI = a(NI, K) + a(TNI, K) * DT, DUMFNS = a(NS, K) + a(TNS, K) * DT, DUMFNR = a(NR, K) + a(TNR, K) * DT;
    double DUMC = a(QC, K) + a(TQC, K) * DT, DUMFNC = a(NC, K) + a(TNC, K) * DT, DUMG = a(QG, K) + a(TQG, K) * DT, DUMFNG = a(NG, K) + a(TNG, K) * DT;
    DUMFNI = mx_(0., DUMFNI); DUMFNS = mx_(0., DUMFNS); DUMFNC = mx_(0., DUMFNC); DUMFNR = mx_(0., DUMFNR); DUMFNG = mx_(0., DUMFNG);
    double DLAMI = 0., DLAMR = 0., DLAMC = 0., DLAMS = 0., DLAMG = 0., PGAM = 0.;
    if (DUMI >= QSMALL) { DLAMI = M25_POW(CONS[12] * DUMFNI / DUMI, F4(1.) / DI); DLAMI = mx_(DLAMI, c.LAMMINI); DLAMI = mn_(DLAMI, c.LAMMAXI); }
    if (DUMR >= QSMALL) { DLAMR = M25_POW(PI * RHOW * DUMFNR / DUMR, third); DLAMR = mx_(DLAMR, c.LAMMINR); DLAMR = mn_(DLAMR, c.LAMMAXR); }
    if (DUMC >= QSMALL) {
      PGAM = F4(0.0005714) * (a(NC, K) / F4(1.E6) * RHOK) + F4(0.2714);
      PGAM = F4(1.) / (PGAM * PGAM) - F4(1.);
      PGAM = mx_(PGAM, 2.); PGAM = mn_(PGAM, 10.);
      DLAMC = M25_POW(CONS[26] * DUMFNC * GAM(PGAM + F4(4.)) / (DUMC * GAM(PGAM + F4(1.))), third);
      const double LAMMIN = (PGAM + F4(1.)) / F4(60.E-6), LAMMAX = (PGAM + F4(1.)) / F4(1.E-6);
      DLAMC = mx_(DLAMC, LAMMIN); DLAMC = mn_(DLAMC, LAMMAX);
    }
    if (DUMQS >= QSMALL) { DLAMS = M25_POW(CONS[1] * DUMFNS / DUMQS, F4(1.) / DS); DLAMS = mx_(DLAMS, c.LAMMINS); DLAMS = mn_(DLAMS, c.LAMMAXS); }
    if (DUMG >= QSMALL) { DLAMG = M25_POW(CONS[2] * DUMFNG / DUMG, F4(1.) / DG); DLAMG = mx_(DLAMG, c.LAMMING); DLAMG = mn_(DLAMG, c.LAMMAXG); }
    const double ACN = a(ACN_, K), AIN = a(AIN_, K), ARN = a(ARN_, K), ASN = a(ASN_, K), AGN = a(AGN_, K);
    double UNC, UMC, UNI, UMI, UNR, UMR, UMS, UNS, UMG, UNG;
    if (DUMC >= QSMALL) {
      UNC = ACN * GAM(F4(1.) + BC + PGAM) / (M25_POW(DLAMC, BC) * GAM(PGAM + F4(1.)));
      UMC = ACN * GAM(F4(4.) + BC + PGAM) / (M25_POW(DLAMC, BC) * GAM(PGAM + F4(4.)));
    } else { UMC = 0.; UNC = 0.; }
    if (DUMI >= QSMALL) { UNI = AIN * CONS[27] / M25_POW(DLAMI, BI); UMI = AIN * CONS[28] / (M25_POW(DLAMI, BI)); } else { UMI = 0.; UNI = 0.; }
    if (DUMR >= QSMALL) { UNR = ARN * CONS[6] / M25_POW(DLAMR, BR); UMR = ARN * CONS[4] / (M25_POW(DLAMR, BR)); } else { UMR = 0.; UNR = 0.; }
    if (DUMQS >= QSMALL) { UMS = ASN * CONS[3] / (M25_POW(DLAMS, BS)); UNS = ASN * CONS[5] / M25_POW(DLAMS, BS); } else { UMS = 0.; UNS = 0.; }
    if (DUMG >= QSMALL) { UMG = AGN * CONS[7] / (M25_POW(DLAMG, BG)); UNG = AGN * CONS[8] / M25_POW(DLAMG, BG); } else { UMG = 0.; UNG = 0.; }
    const double dum = M25_POW(RHOSU / RHOK, F4(0.54));
    UMS = mn_(UMS, F4(1.2) * dum); UNS = mn_(UNS, F4(1.2) * dum);
    UMI = mn_(UMI, F4(1.2) * M25_POW(RHOSU / RHOK, F4(0.35))); UNI = mn_(UNI, F4(1.2) * M25_POW(RHOSU / RHOK, F4(0.35)));
    UMR = mn_(UMR, F4(9.1) * dum); UNR = mn_(UNR, F4(9.1) * dum);
    UMG = mn_(UMG, F4(20.) * dum); UNG = mn_(UNG, F4(20.) * dum);
    double fr = UMR, fi = UMI, fni = UNI, fs = UMS, fns = UNS, fnr = UNR, fc = UMC, fnc = UNC, fg = UMG, fng = UNG;
    if (K <= nz - 1) {
      if (fr < F4(1.E-10)) fr = frn;
      if (fi < F4(1.E-10)) fi = fin;
      if (fni < F4(1.E-10)) fni = fnin;
      if (fs < F4(1.E-10)) fs = fsn;
      if (fns < F4(1.E-10)) fns = fnsn;
      if (fnr < F4(1.E-10)) fnr = fnrn;
      if (fc < F4(1.E-10)) fc = fcn;
      if (fnc < F4(1.E-10)) fnc = fncn;
      if (fg < F4(1.E-10)) fg = fgn;
      if (fng < F4(1.E-10)) fng = fngn;
    }
    frn = fr; fin = fi; fnin = fni; fsn = fs; fnsn = fns; fnrn = fnr; fcn = fc; fncn = fnc; fgn = fg; fngn = fng;
    a(FR, K) = fr; a(FI, K) = fi; a(FNI, K) = fni; a(FS, K) = fs; a(FNS, K) = fns; a(FNR, K) = fnr; a(FC, K) = fc; a(FNC, K) = fnc; a(FG, K) = fg; a(FNG, K) = fng;
    const double RGVM = mx_(mx_(mx_(mx_(mx_(mx_(mx_(mx_(mx_(fr, fi), fs), fc), fni), fnr), fns), fnc), fg), fng);
    const int ns = (int)(RGVM * DT / a(DZQ, K) + F4(1.));
    if (ns > NSTEP) NSTEP = ns;
    a(DR, K) = DUMR * RHOK; a(DI_, K) = DUMI * RHOK; a(DNI, K) = DUMFNI * RHOK; a(DS_, K) = DUMQS * RHOK; a(DNS, K) = DUMFNS * RHOK;
    a(DNR, K) = DUMFNR * RHOK; a(DC, K) = DUMC * RHOK; a(DNC, K) = DUMFNC * RHOK; a(DG_, K) = DUMG * RHOK; a(DNG, K) = DUMFNG * RHOK;
  }
  const double rn = (double)NSTEP;
  const int DID[10] = {DR, DI_, DNI, DS_, DNS, DNR, DC, DNC, DG_, DNG};
  const int FID[10] = {FR, FI, FNI, FS, FNS, FNR, FC, FNC, FG, FNG};
  // species order of DID: mass tendencies -> SQ* (+ nothing else), number tendencies -> *3DTEN (the N*STEN arrays of the reference feed nothing)
  const int TID[10] = {SQR, SQI, TNI, SQNI, TNS, TNR, SQC, TNC, SQG, TNG};
#if defined(__CUDA_ARCH__) || defined(CRM_EMU)
  // Device order of the same arithmetic: species outermost, then levels from the top, then the NSTEP sub-steps of that level with the
  // level's density and tendency in registers (the fallout of the level above is kept per sub-step: fprev[n]).  Every (species, level,
  // sub-step) update is the reference's expression on the same operands, executed in another order of INDEPENDENT updates, so the
  // results are bit-identical; the work space is read and written once per species and level instead of once per sub-step.
  constexpr int NMAX = 8;
  if (NSTEP <= NMAX) {
    double f1[5][NMAX];   // fallout at the surface of rain, cloud water, snow, cloud ice, graupel (the terms of PRECRT / SNOWRT)
    for (int s = 0; s < 10; s++) {
      // fallout of the level above, per sub-step.  Starting from +0 makes the top level the general case: (0 - fout)/dzq = -(fout/dzq)
      // and T + (-x) = T - x exactly, which is the reference's K = NZ branch.
      double fprev[NMAX];
#pragma unroll
      for (int n = 0; n < NMAX; n++) fprev[n] = 0.;
      const int slot = (s == 0) ? 0 : (s == 6) ? 1 : (s == 3) ? 2 : (s == 1) ? 3 : (s == 8) ? 4 : -1;
      for (int K = nz; K >= 1; K--) {
        const double dzq = a(DZQ, K), RHOK = a(RHO, K), Fk = a(FID[s], K);
        double Dk = a(DID[s], K), Tk = a(TID[s], K);
#pragma unroll
        for (int n = 0; n < NMAX; n++) {
          if (n < NSTEP) {
            const double fout = Fk * Dk;
            const double FALTND = (fprev[n] - fout) / dzq;
            Tk = Tk + FALTND / rn / RHOK; Dk = Dk + FALTND * DT / rn;
            fprev[n] = fout;
          }
        }
        a(TID[s], K) = Tk; a(DID[s], K) = Dk;
      }
      if (slot >= 0) {
#pragma unroll
        for (int n = 0; n < NMAX; n++) f1[slot][n] = fprev[n];   // after the last level fprev holds the fallout of level 1
      }
    }
    for (int n = 0; n < NSTEP; n++) {
      PRECRT = PRECRT + (f1[0][n] + f1[1][n] + f1[2][n] + f1[3][n] + f1[4][n]) * DT / rn;
      SNOWRT = SNOWRT + (f1[2][n] + f1[3][n] + f1[4][n]) * DT / rn;
    }
  } else
#endif
  for (int N = 1; N <= NSTEP; N++) {
    double fout_up[10];   // FALOUT*(K+1) of this sub-step (values before the update of level K+1)
    double fout1[10];
    for (int K = nz; K >= 1; K--) {
      const double dzq = a(DZQ, K), RHOK = a(RHO, K);
      for (int s = 0; s < 10; s++) {
        const double fout = a(FID[s], K) * a(DID[s], K);
        double FALTND;
        if (K == nz) {
          FALTND = fout / dzq;
          a(TID[s], K) = a(TID[s], K) - FALTND / rn / RHOK;
          a(DID[s], K) = a(DID[s], K) - FALTND * DT / rn;
        } else {
          FALTND = (fout_up[s] - fout) / dzq;
          a(TID[s], K) = a(TID[s], K) + FALTND / rn / RHOK;
          a(DID[s], K) = a(DID[s], K) + FALTND * DT / rn;
        }
        fout_up[s] = fout;
        if (K == 1) fout1[s] = fout;
      }
    }
    PRECRT = PRECRT + (fout1[0] + fout1[6] + fout1[3] + fout1[1] + fout1[8]) * DT / rn;
    SNOWRT = SNOWRT + (fout1[3] + fout1[1] + fout1[8]) * DT / rn;
  }

  if (PH == 2) { a(W3, 2) = PRECRT; a(W3, 3) = SNOWRT; return; }
  }   // PH 0 / 2
  // ------------------------------------------------------------------- final update (5129-5604)
  for (int K = 1; K <= nz; K++) {
    const double PRESK = a(PRES, K), RHOK = a(RHO, K);
    double QR3DTEN = a(TQR, K) + a(SQR, K), QI3DTEN = a(TQI, K) + a(SQI, K), QC3DTEN = a(TQC, K) + a(SQC, K), QG3DTEN = a(TQG, K) + a(SQG, K),
           QNI3DTEN = a(TQNI, K) + a(SQNI, K);
    double NS3DTEN = a(TNS, K), NI3DTEN = a(TNI, K);
    double QC3D = a(QC, K), QI3D = a(QI, K), QNI3D = a(QNI, K), QR3D = a(QR, K), NC3D = a(NC, K), NI3D = a(NI, K), NS3D = a(NS, K), NR3D = a(NR, K),
           QG3D = a(QG, K), NG3D = a(NG, K), T3D = a(T3, K), QV3D = a(QV, K);
    double mlt1d = a(MLT, K);
    double LAMI = a(LAMI_, K);
    if (QI3D >= QSMALL && T3D < TMELT && LAMI >= F4(1.E-10)) {
      if (F4(1.) / LAMI >= F4(2.) * DCS) {
        QNI3DTEN = QNI3DTEN + QI3D / DT + QI3DTEN;
        NS3DTEN = NS3DTEN + NI3D / DT + NI3DTEN;
        QI3DTEN = -QI3D / DT;
        NI3DTEN = -NI3D / DT;
      }
    }
    QC3D = QC3D + QC3DTEN * DT; QI3D = QI3D + QI3DTEN * DT; QNI3D = QNI3D + QNI3DTEN * DT; QR3D = QR3D + QR3DTEN * DT;
    NC3D = NC3D + a(TNC, K) * DT; NI3D = NI3D + NI3DTEN * DT; NS3D = NS3D + NS3DTEN * DT; NR3D = NR3D + a(TNR, K) * DT;
    QG3D = QG3D + QG3DTEN * DT; NG3D = NG3D + a(TNG, K) * DT;
    T3D = T3D + a(TT, K) * DT;
    QV3D = QV3D + a(TQV, K) * DT;
    double EVS = mn_(F4(0.99) * PRESK, polysvp(T3D, 0));
    double EIS = mn_(F4(0.99) * PRESK, polysvp(T3D, 1));
    if (EIS > EVS) EIS = EVS;
    const double QVS = EP_2 * EVS / (PRESK - EVS);
    const double QVI = EP_2 * EIS / (PRESK - EIS);
    const double QVQVS = QV3D / QVS, QVQVSI = QV3D / QVI;
    if (QVQVS < F4(0.9)) {
      if (QR3D < F4(1.E-8)) { QV3D = QV3D + QR3D; T3D = T3D - QR3D * XXLV / CPM; QR3D = 0.; }
      if (QC3D < F4(1.E-8)) { QV3D = QV3D + QC3D; T3D = T3D - QC3D * XXLV / CPM; QC3D = 0.; }
    }
    if (QVQVSI < F4(0.9)) {
      if (QI3D < F4(1.E-8)) { QV3D = QV3D + QI3D; T3D = T3D - QI3D * XXLS / CPM; QI3D = 0.; }
      if (QNI3D < F4(1.E-8)) { QV3D = QV3D + QNI3D; T3D = T3D - QNI3D * XXLS / CPM; QNI3D = 0.; }
      if (QG3D < F4(1.E-8)) { QV3D = QV3D + QG3D; T3D = T3D - QG3D * XXLS / CPM; QG3D = 0.; }
    }
    if (QC3D < QSMALL) { QV3D = QV3D + QC3D; T3D = T3D - QC3D * XXLV / CPM; QC3D = 0.; NC3D = 0.; }
    if (QR3D < QSMALL) { QV3D = QV3D + QR3D; T3D = T3D - QR3D * XXLV / CPM; QR3D = 0.; NR3D = 0.; }
    if (QI3D < QSMALL) { QV3D = QV3D + QI3D; T3D = T3D - QI3D * XXLS / CPM; QI3D = 0.; NI3D = 0.; }
    if (QNI3D < QSMALL) { QV3D = QV3D + QNI3D; T3D = T3D - QNI3D * XXLS / CPM; QNI3D = 0.; NS3D = 0.; }
    if (QG3D < QSMALL) { QV3D = QV3D + QG3D; T3D = T3D - QG3D * XXLS / CPM; QG3D = 0.; NG3D = 0.; }
    const bool skip500 = (QC3D < QSMALL && QI3D < QSMALL && QNI3D < QSMALL && QR3D < QSMALL && QG3D < QSMALL);
    if (!skip500) {
      if (QI3D >= QSMALL && T3D >= TMELT) {   // instantaneous melting of cloud ice
        QR3D = QR3D + QI3D; T3D = T3D - QI3D * XLF / CPM;
        mlt1d = mlt1d + QI3D / DT;
        QI3D = 0.; NR3D = NR3D + NI3D; NI3D = 0.;
      }
      if (T3D <= F4(233.15) && QC3D >= QSMALL) {   // homogeneous freezing
        QI3D = QI3D + QC3D; T3D = T3D + QC3D * XLF / CPM; QC3D = 0.; NI3D = NI3D + NC3D; NC3D = 0.;
      }
      if (T3D <= F4(233.15) && QR3D >= QSMALL) { QG3D = QG3D + QR3D; T3D = T3D + QR3D * XLF / CPM; QR3D = 0.; NG3D = NG3D + NR3D; NR3D = 0.; }
      NI3D = mx_(0., NI3D); NS3D = mx_(0., NS3D); NC3D = mx_(0., NC3D); NR3D = mx_(0., NR3D); NG3D = mx_(0., NG3D);
      if (QI3D >= QSMALL) {
        LAMI = M25_POW(CONS[12] * NI3D / QI3D, F4(1.) / DI);
        if (LAMI < c.LAMMINI) { LAMI = c.LAMMINI; const double N0I = M25_POW(LAMI, DI + F4(1.)) * QI3D / CONS[12]; NI3D = N0I / LAMI; }
        else if (LAMI > c.LAMMAXI) { LAMI = c.LAMMAXI; const double N0I = M25_POW(LAMI, DI + F4(1.)) * QI3D / CONS[12]; NI3D = N0I / LAMI; }
      }
      if (QR3D >= QSMALL) {
        double LAMR = M25_POW(PI * RHOW * NR3D / QR3D, third);
        if (LAMR < c.LAMMINR) { LAMR = c.LAMMINR; const double N0RR = p4_(LAMR) * QR3D / (PI * RHOW); NR3D = N0RR / LAMR; }
        else if (LAMR > c.LAMMAXR) { LAMR = c.LAMMAXR; const double N0RR = p4_(LAMR) * QR3D / (PI * RHOW); NR3D = N0RR / LAMR; }
      }
      if (QC3D >= QSMALL) {
        double PGAM = F4(0.0005714) * (NC3D / F4(1.E6) * RHOK) + F4(0.2714);
        PGAM = F4(1.) / (PGAM * PGAM) - F4(1.);
        PGAM = mx_(PGAM, 2.); PGAM = mn_(PGAM, 10.);
        double LAMC = M25_POW(CONS[26] * NC3D * GAM(PGAM + F4(4.)) / (QC3D * GAM(PGAM + F4(1.))), third);
        const double LAMMIN = (PGAM + F4(1.)) / F4(60.E-6), LAMMAX = (PGAM + F4(1.)) / F4(1.E-6);
        if (LAMC < LAMMIN) { LAMC = LAMMIN; NC3D = M25_EXP(F4(3.) * M25_LOG(LAMC) + M25_LOG(QC3D) + M25_LOG(GAM(PGAM + F4(1.))) - M25_LOG(GAM(PGAM + F4(4.)))) / CONS[26]; }
        else if (LAMC > LAMMAX) { LAMC = LAMMAX; NC3D = M25_EXP(F4(3.) * M25_LOG(LAMC) + M25_LOG(QC3D) + M25_LOG(GAM(PGAM + F4(1.))) - M25_LOG(GAM(PGAM + F4(4.)))) / CONS[26]; }
      }
      if (QNI3D >= QSMALL) {
        double LAMS = M25_POW(CONS[1] * NS3D / QNI3D, F4(1.) / DS);
        if (LAMS < c.LAMMINS) { LAMS = c.LAMMINS; const double N0S = M25_POW(LAMS, DS + F4(1.)) * QNI3D / CONS[1]; NS3D = N0S / LAMS; }
        else if (LAMS > c.LAMMAXS) { LAMS = c.LAMMAXS; const double N0S = M25_POW(LAMS, DS + F4(1.)) * QNI3D / CONS[1]; NS3D = N0S / LAMS; }
      }
      if (QG3D >= QSMALL) {
        double LAMG = M25_POW(CONS[2] * NG3D / QG3D, F4(1.) / DG);
        if (LAMG < c.LAMMING) { LAMG = c.LAMMING; const double N0G = M25_POW(LAMG, DG + F4(1.)) * QG3D / CONS[2]; NG3D = N0G / LAMG; }
        else if (LAMG > c.LAMMAXG) { LAMG = c.LAMMAXG; const double N0G = M25_POW(LAMG, DG + F4(1.)) * QG3D / CONS[2]; NG3D = N0G / LAMG; }
      }
    }
    // 500 CONTINUE
    NI3D = mn_(NI3D, F4(10.E6) / RHOK);
    NC3D = mn_(NC3D, (c.NANEW1 + c.NANEW2) / RHOK);   // INUM = 0 and IACT = 2
    a(QC, K) = QC3D; a(QI, K) = QI3D; a(QNI, K) = QNI3D; a(QR, K) = QR3D; a(NC, K) = NC3D; a(NI, K) = NI3D; a(NS, K) = NS3D; a(NR, K) = NR3D;
    a(QG, K) = QG3D; a(NG, K) = NG3D; a(T3, K) = T3D; a(QV, K) = QV3D; a(MLT, K) = mlt1d;
  }
}
template <class A, class Cn>
M25_HD inline void column(A &a, const Cn &c, const int nz, const double DT, double &PRECRT, double &SNOWRT) {
  column_ph<0, false>(a, c, nz, DT, PRECRT, SNOWRT);
}

}  // namespace m2005
