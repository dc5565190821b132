// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
// M2005 two-moment microphysics of the CRM for ONE CRM: restatement of micro_proc / micro_diagnose of
// components/cam/src/physics/crm/MICRO_M2005/microphysics.F90 (:636-1297, :1305-1345) around the column scheme
// M2005MICRO_GRAUPEL (module_mp_graupel.F90:1102-5610).
// The column scheme itself is NOT restated a second time here: this file compiles the single transcription in
// acme-ecp_b200/csrc/crm_m2005_col.h (plain C++, also compiled by nvcc into the product kernel) with g++ -O2
// -ffp-contract=off.  GPU-vs-oracle parity of the m2005 configuration therefore checks the driver, the memory plumbing and
// the compilers, not the transcription of the 2 800 lines of column physics; that transcription is pinned by
// tests/test_m2005_kat.py only.
#include "crm_state.h"
#include "crm_routines.h"
#include "../acme-ecp_b200/csrc/crm_m2005_col.h"

namespace orc {

static m2005::Const g_m25;
static bool g_m25_init = false;
const m2005::Const &m2005_const() {
  if (!g_m25_init) { m2005::graupel_init(g_m25, rgas, rv, cp, lcond, lsub); g_m25_init = true; }
  return g_m25;
}

// micro_diagnose (MICRO_M2005/microphysics.F90:1305-1345, non-CLUBB): qv = qt - cloudliq, qcl = cloudliq, qpl = qr,
// qci = qi, qpi = qs + qg.  cloudliq lives in c.qn.
void micro_diagnose_m2005(const Geo &g, Crm &c) {
  for (int k = 1; k <= g.nzm; k++)
    for (int j = 1; j <= g.ny; j++)
      for (int i = 1; i <= g.nx; i++) {
        c.qv(i, j, k) = c.qt(i, j, k) - c.qn(i, j, k);
        c.qcl(i, j, k) = c.qn(i, j, k);
        c.qpl(i, j, k) = c.mf[3](i, j, k);
        c.qci(i, j, k) = c.mf[5](i, j, k);
        c.qpi(i, j, k) = c.mf[7](i, j, k) + c.mf[9](i, j, k);
      }
}

struct ColArr {
  std::vector<double> d; int nz;
  ColArr(int nz_) : d((size_t)m2005::NARR * nz_, 0.), nz(nz_) {}
  inline double &operator()(int id, int k) { return d[(size_t)id * nz + (k - 1)]; }
};

// micro_proc (MICRO_M2005/microphysics.F90:636-1297) with dostatis = .false. (crm_module.F90:295), doprecip, docloud,
// doicemicro, dograupel, dopredictNc = .true.
void micro_proc_m2005(const Geo &g, Crm &c, const Step &s) {
  using namespace m2005;
  const m2005::Const &cc = m2005_const();
  const int nx = g.nx, ny = g.ny, nzm = g.nzm, nz = g.nz;
  const double dtn = s.dtn;
  F3 &qtf = c.qt;
  std::vector<double> stend_qv(nzm + 1, 0.), stend_qr(nzm + 1, 0.), stend_qci(nzm + 1, 0.), stend_qs(nzm + 1, 0.), stend_qg(nzm + 1, 0.);
  ColArr a(nzm);
  for (int j = 1; j <= ny; j++)
    for (int i = 1; i <= nx; i++) {
      std::fill(a.d.begin(), a.d.end(), 0.);
      std::vector<double> tmpqcl(nzm + 1, 0.);
      for (int k = 1; k <= nzm; k++) {
        a(QV, k) = qtf(i, j, k);           // total water qv + qcl
        a(NC, k) = c.mf[2](i, j, k); a(QR, k) = c.mf[3](i, j, k); a(NR, k) = c.mf[4](i, j, k); a(QI, k) = c.mf[5](i, j, k);
        a(NI, k) = c.mf[6](i, j, k); a(QNI, k) = c.mf[7](i, j, k); a(NS, k) = c.mf[8](i, j, k); a(QG, k) = c.mf[9](i, j, k); a(NG, k) = c.mf[10](i, j, k);
        a(T3, k) = c.t(i, j, k) - c.gamaz(k) + fac_cond * (a(QR, k)) + fac_sub * (a(QI, k) + a(QNI, k) + a(QG, k));
        a(DZQ, k) = c.adz(k) * c.dz;
        a(W3, k) = ((c.zi(k + 1) - c.z(k)) * c.w(i, j, k) + (c.z(k) - c.zi(k)) * c.w(i, j, k + 1)) / (c.zi(k + 1) - c.zi(k));
        a(WVAR, k) = std::sqrt(c.tke(i, j, k) / R4(3.));
        c.wvar(i, j, k) = a(WVAR, k);
        a(PRES, k) = R4(100.) * c.pres(k);
        a(RHO, k) = c.rho(k);
        satadj_liquid(cc, a(T3, k), a(QV, k), a(QC, k), a(PRES, k));
      }
      double sfcpcp = 0., sfcicepcp = 0.;
      m2005::column(a, cc, nzm, dtn, sfcpcp, sfcicepcp);
      for (int k = 1; k <= nzm; k++) {
        c.aut1(i, j, k) = a(AUT, k); c.acc1(i, j, k) = a(ACC, k); c.evpc1(i, j, k) = a(EVPC, k); c.evpr1(i, j, k) = a(EVPR, k);
        c.mlt1(i, j, k) = a(MLT, k); c.sub1(i, j, k) = a(SUB, k); c.dep1(i, j, k) = a(DEP, k); c.con1(i, j, k) = a(CON, k);
        c.aut1a(i, j, k) += c.aut1(i, j, k) * dtn; c.acc1a(i, j, k) += c.acc1(i, j, k) * dtn; c.evpc1a(i, j, k) += c.evpc1(i, j, k) * dtn;
        c.evpr1a(i, j, k) += c.evpr1(i, j, k) * dtn; c.mlt1a(i, j, k) += c.mlt1(i, j, k) * dtn; c.sub1a(i, j, k) += c.sub1(i, j, k) * dtn;
        c.dep1a(i, j, k) += c.dep1(i, j, k) * dtn; c.con1a(i, j, k) += c.con1(i, j, k) * dtn;
      }
      const int o2 = (j - 1) * nx + i - 1;
      c.precsfc[o2] = c.precsfc[o2] + sfcpcp / c.dz;
      c.precssfc[o2] = c.precssfc[o2] + sfcicepcp / c.dz;
      for (int k = 1; k <= nzm; k++) {
        c.mf[3](i, j, k) = a(QR, k); c.mf[4](i, j, k) = a(NR, k);
        qtf(i, j, k) = a(QV, k) + a(QC, k);
        c.qn(i, j, k) = a(QC, k);
        c.mf[2](i, j, k) = a(NC, k);
        c.mf[5](i, j, k) = a(QI, k); c.mf[6](i, j, k) = a(NI, k); c.mf[7](i, j, k) = a(QNI, k); c.mf[8](i, j, k) = a(NS, k);
        c.mf[9](i, j, k) = a(QG, k); c.mf[10](i, j, k) = a(NG, k);
        c.t(i, j, k) = c.t(i, j, k) - dtn * fac_cond * (a(SQC, k) + a(SQR, k)) - dtn * fac_sub * (a(SQI, k) + a(SQNI, k) + a(SQG, k));
        stend_qv[k] += a(SQC, k); stend_qr[k] += a(SQR, k); stend_qci[k] += a(SQI, k); stend_qs[k] += a(SQNI, k); stend_qg[k] += a(SQG, k);
      }
    }
  // sedimentation fluxes (microphysics.F90:1202-1258)
  std::vector<double> mk(nzm + 1);
  auto add_flux = [&](const std::vector<double> &st) {
    double tmp = 0.;
    for (int k = 1; k <= nzm; k++) {
      int m = nz - k;
      tmp = tmp + st[m] * c.rho(m) * c.dz * c.adz(m);
      mk[m] = tmp;
    }
  };
  add_flux(stend_qv);
  for (int k = 1; k <= nzm; k++) c.precflux(k) = c.precflux(k) - mk[k] * dtn / c.dz;
  add_flux(stend_qr);
  for (int k = 1; k <= nzm; k++) c.precflux(k) = c.precflux(k) - mk[k] * dtn / c.dz;
  std::vector<double> mi(nzm + 1), ms(nzm + 1), mg(nzm + 1);
  add_flux(stend_qci); mi = mk; add_flux(stend_qs); ms = mk; add_flux(stend_qg); mg = mk;
  for (int k = 1; k <= nzm; k++) c.precflux(k) = c.precflux(k) - (mi[k] + ms[k] + mg[k]) * dtn / c.dz;
  micro_diagnose_m2005(g, c);
}

}  // namespace orc

// ---- test hooks: the REAL(4)-valued helper functions and the column scheme on caller-supplied columns ----
extern "C" double crm_oracle_m2005_polysvp(double t, int type) { return m2005::polysvp(t, type); }
extern "C" double crm_oracle_m2005_gamma(double x) { const m2005::Const &c = orc::m2005_const(); return m2005::gamma_f(x, c.PI, c.SQRTPI); }
extern "C" double crm_oracle_m2005_derf1(double x) { return m2005::derf1(x); }
extern "C" double crm_oracle_m2005_const(int which) {
  const m2005::Const &c = orc::m2005_const();
  if (which >= 1 && which <= 41) return c.CONS[which];
  switch (which) { case 100: return c.MI0; case 101: return c.MMULT; case 102: return c.RHOSU; case 103: return c.BACT; case 104: return c.F11; case 105: return c.LAMMINI;
                   case 106: return c.LAMMAXR; default: return 0.; }
}
extern "C" void crm_oracle_m2005_satadj(double *tabs, double *qt, double *qc, double pres) { m2005::satadj_liquid(orc::m2005_const(), *tabs, *qt, *qc, pres); }
// state: (12, nz) rows QC,QI,QNI,QR,NC,NI,NS,NR,QG,NG,T,QV; inp: (5, nz) rows PRES,RHO,DZQ,W,WVAR; sed: (5, nz) QRSTEN,QISTEN,QNISTEN,QCSTEN,QGSTEN;
// rates: (8, nz); prec: [PRECRT, SNOWRT]
extern "C" void crm_oracle_m2005_column(int nz, double dt, double *state, const double *inp, double *sed, double *rates, double *prec) {
  using namespace m2005;
  orc::ColArr a(nz);
  for (int id = 0; id < 12; id++) for (int k = 1; k <= nz; k++) a(id, k) = state[id * nz + k - 1];
  for (int q = 0; q < 5; q++) for (int k = 1; k <= nz; k++) a(PRES + q, k) = inp[q * nz + k - 1];
  m2005::column(a, orc::m2005_const(), nz, dt, prec[0], prec[1]);
  for (int id = 0; id < 12; id++) for (int k = 1; k <= nz; k++) state[id * nz + k - 1] = a(id, k);
  for (int q = 0; q < 5; q++) for (int k = 1; k <= nz; k++) sed[q * nz + k - 1] = a(SQR + q, k);
  for (int q = 0; q < 8; q++) for (int k = 1; k <= nz; k++) rates[q * nz + k - 1] = a(AUT + q, k);
}
