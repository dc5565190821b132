// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).
// Per-CRM state: the module globals of CRM/vars.F90, grid.F90, params.F90, SGS_TKE/sgs.F90,
// MICRO_SAM1MOM/microphysics.F90 + micro_params.F90 for ONE CRM (the reference stores the batch index
// icrm fastest; CRMs never interact except through the batch-global maxima handled by the driver).
#pragma once
#include "oracle_common.h"

namespace orc {

struct Geo {   // PHYS/crmdims.F90, CRM/domain.F90, CRM/grid.F90:14-46
  int nx, ny, nzm, nz, plev, YES3D;
  int nx_rad, ny_rad;
  double dx, dy;
  bool RUN3D;
  bool m2005 = false;   // micro = 1
  bool esmt = false;    // -DSP_ESMT
};

// Scalars shared by the whole batch (grid.F90:52-60, 101, 181)
struct Step {
  int nstep = 0, ncycle = 1, icycle = 1, na = 1, nb = 2, nc = 3, nstop = 0;
  double at = 0, bt = 0, ct = 0, dtn = 0, dtfactor = 0, dt = 0;
  double dt3[4] = {0, 0, 0, 0};   // 1-based
};

struct Crm {
  // ---- grid.F90:176-183, vars.F90:219-227 ----
  P1 z, zi, pres, presi, adz, adzw, rho, rhow, bet, gamaz;
  double dz = 0;
  // ---- prognostic / diagnostic 3-D fields (vars.F90:172-188, App. B of SURVEY.md) ----
  F3 u, v, w, t, p, tabs, qv, qcl, qpl, qci, qpi;
  F3 dudt[4], dvdt[4], dwdt[4];   // 1-based AB slot
  F3 cf3d;
  // microphysics.F90:27-50 (sam1mom): micro_field(:,:,:,:,1)=qt, (..,2)=qp
  F3 qt, qp, qn;
  // MICRO_M2005/microphysics.F90:57-121 (micro = 1): micro_field(:,:,:,:,n) = mf[n], n = 2..10 (nc, qr, nr, qi, ni, qs, ns, qg, ng);
  // micro_field(..,1) = qt (total water vapour + cloud liquid); cloudliq lives in qn; qp is unused
  F3 mf[11];
  F3 u_esmt, v_esmt;   // scalar_momentum.F90:27-28 (SP_ESMT)
  F3 wvar, aut1, acc1, evpc1, evpr1, mlt1, sub1, dep1, con1, aut1a, acc1a, evpc1a, evpr1a, mlt1a, sub1a, dep1a, con1a;
  F3 rad_nc, rad_ni, rad_qs, rad_ns;
  // sgs.F90:40-54
  F3 tke, tk, tkh;
  P1 grdf_x, grdf_y, grdf_z;
  // 2-D
  std::vector<double> precsfc, precssfc, fluxbu, fluxbv, fluxbt, fluxbq, fluxtu, fluxtv, fluxtt, fluxtq;
  // profiles (vars.F90:200-253)
  P1 t0, q0, qv0, tabs0, u0, v0, tg0, qg0, ug0, vg0, tke0, t01, q01, qp0, qn0;
  P1 qtend, ttend, utend, vtend;
  P1 precflux, uwle, uwsb, vwle, vwsb, tlat, tlatqi, qifall, qpfall, twsb, tdiff;
  P1 mkwle[11], mkwsb[11], mkadv[11], mkdiff[11];   // index = micro field (1,2 sam1mom; 1..10 m2005)
  P1 qpsrc, qpevp;
  // micro_params.F90:82-93
  P1 accrsc, accrsi, accrrc, coefice, accrgc, accrgi, evaps1, evaps2, evapr1, evapr2, evapg1, evapg2;
  // params.F90
  double z0 = R4(0.035), uhl = 0, vhl = 0, taux0 = 0, tauy0 = 0;
  // crm() locals that persist over the call (crm_module.F90:142-180)
  P1 t00, tln, qln, qccln, qiiln, uln, vln, dd_crm, mui_crm, mdi_crm;
  double colprec = 0, colprecs = 0, bflx = 0, wnd = 0, ustar = 0;
  double prectend0 = 0, precstend0 = 0;
  // ice_fall / kurant scratch
  int kmin = 0, kmax = 0;
  // precip_fall scratch that must survive the batch-global nprec reduction
  F3 pf_wp, pf_lfac, pf_omega;
  // accelerate_crm scratch
  P1 acc_t, acc_q, acc_u, acc_v;
  // crm_output accumulators, GCM-level indexed (1..plev)
  P1 o_cld, o_cldtop, o_gicewp, o_gliqwp, o_mcup, o_mcdn, o_mcuup, o_mcudn;
  double o_cltot = 0, o_clhgh = 0, o_clmed = 0, o_cllow = 0, o_timing = 0;
  // crm_rad accumulators (nx_rad,ny_rad,nzm)
  F3 rad_t, rad_qv, rad_qc, rad_qi, rad_cld;
};

}  // namespace orc
