// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle_common.h).  Routine prototypes (one CRM each).
#pragma once
#include "crm_state.h"

namespace orc {
// crm_dycore.cpp
void bound_exchange(const Geo &g, F3 &f, int nk, int i_1, int i_2, int j_1, int j_2);
void boundaries(const Geo &g, Crm &c, int flag);
void abcoefs(Step &s);
void zero(const Geo &g, Crm &c, const Step &s);
void buoyancy(const Geo &g, Crm &c, const Step &s);
void forcing(const Geo &g, Crm &c, const Step &s);
void rad_heating(const Geo &g, Crm &c, const Step &s, const F3 &qrad);
int damping(const Geo &g, Crm &c, const Step &s);
void ice_fall(const Geo &g, Crm &c, const Step &s);
void sgs_proc(const Geo &g, Crm &c, const Step &s);
void advect_mom(const Geo &g, Crm &c, const Step &s);
void sgs_mom(const Geo &g, Crm &c, const Step &s);
void pressure(const Geo &g, Crm &c, const Step &s, const FftPlan &fx, const FftPlan &fy);
void adams(const Geo &g, Crm &c, const Step &s);
void uvw(const Geo &g, Crm &c, const Step &s);
void advect_scalar(const Geo &g, const Crm &c, F3 &f, P1 &fadv, P1 &flux);
void diffuse_scalar(const Geo &g, const Crm &c, const Step &s, F3 &f, const std::vector<double> &fluxb,
                    const std::vector<double> &fluxt, P1 &fdiff, P1 &flux);
double kurant_cfl(const Geo &g, const Crm &c, const Step &s);
// crm_esmt.cpp (-DSP_ESMT)
void esmt_advect(const Geo &g, Crm &c);
void esmt_diffuse(const Geo &g, Crm &c, const Step &s);
void scalar_momentum_tend(const Geo &g, Crm &c, const Step &s);
// crm_micro.cpp
struct MicroCoef {   // micro_params.F90:70-97 + microphysics.F90:32 scalars
  float gam3, gams1, gams2, gams3, gamg1, gamg2, gamg3, gamr1, gamr2, gamr3;
  double a_bg, a_pr, a_gr;
  double vrain, vsnow, vgrau, crain, csnow, cgrau;
  void init();
};
void micro_diagnose(const Geo &g, Crm &c, const MicroCoef &m);
void precip_init(const Geo &g, Crm &c, const MicroCoef &m);
void cloud(const Geo &g, Crm &c);
void precip_proc(const Geo &g, Crm &c, const Step &s, const MicroCoef &m);
double precip_fall_part1(const Geo &g, Crm &c, const Step &s, const MicroCoef &m);   // returns local prec_cfl
void precip_fall_part2(const Geo &g, Crm &c, const Step &s, const MicroCoef &m, double prec_cfl);
void diagnose(const Geo &g, Crm &c, const Step &s);
bool accelerate_part1(const Geo &g, Crm &c, bool accel_uv);   // returns local ceaseflag
void accelerate_part2(const Geo &g, Crm &c, bool accel_uv, double accel_factor);
void setperturb(const Geo &g, Crm &c, int iseed);
double z0_est(double z, double bflx, double wnd, double ustar);
// crm_m2005.cpp
void micro_diagnose_m2005(const Geo &g, Crm &c);
void micro_proc_m2005(const Geo &g, Crm &c, const Step &s);
}  // namespace orc
