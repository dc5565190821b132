/*
 * crm_b200.h -- C ABI of the B200-native embedded cloud-resolving model (CRM).
 *
 * This is the drop-in boundary for the one call site of the reference's CRM driver,
 *
 *     call crm(lchnk, icol, ncrms, dt_gl, plev, crm_input, crm_state, crm_rad,
 *              crm_ecpp_output, crm_output)
 *
 * (reference: components/cam/src/physics/cam/crm_physics.F90:1145-1156 calling
 *  components/cam/src/physics/crm/crm_module.F90:45-56).  A Fortran ISO_C_BINDING shim
 * (shim/crm_module_b200.F90) does c_loc() on every member array of the reference's
 * derived types and passes the hidden module inputs explicitly.
 *
 * Conventions
 *  - plain C, no C++/torch types; every array is FP64 (real(crm_rknd)/real(r8), reference
 *    CRM/params.F90:21-26) or int32.
 *  - arrays are column-major exactly as the Fortran caller owns them: the GCM-column (CRM)
 *    index is FASTEST with leading dimension `ld` (= pcols in the reference), i.e.
 *    A(icrm,i,j,k) lives at  A[icrm + ld*(i + nx*(j + ny*k))]   (0-based),
 *    profiles P(icrm,l) at P[icrm + ld*l]; only rows 0..ncrms-1 are read/written.
 *  - pointers may be host (pageable or pinned) or device pointers; the library detects which
 *    (cudaPointerGetAttributes).  Ownership stays with the caller.
 *  - outputs are valid on return (synchronous semantics, like the reference).
 *  - a NULL output pointer means "caller does not want this member".
 *  - return value 0 = success; otherwise crm_b200_last_error() describes the failure
 *    (replaces the reference's task_abort/endrun: CRM/kurant.F90:66-75, damping.F90:39-42,
 *    crm_module.F90:690-691).
 */
#ifndef CRM_B200_H
#define CRM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRM_B200_ABI_VERSION 5

/* Compile-time constants of the reference (PHYS/crmdims.F90:13-22, CRM/domain.F90:9-13) and the
 * hidden module inputs of crm() (time_manager:get_nstep, accelerate_crm_mod namelist). */
typedef struct crm_dims_t {
  int ncrms;            /* number of CRMs (GCM columns) in this call           */
  int ld;               /* leading dimension of every array (reference: pcols)   */
  int plev;             /* GCM levels                                            */
  int nx, ny, nzm;      /* crm_nx, crm_ny, crm_nz (scalar levels)                */
  int nx_rad, ny_rad;   /* crm_nx_rad, crm_ny_rad                                */
  int yes3d;            /* YES3DVAL: 1 if ny>1                                   */
  double dx;            /* CRM_DX [m] (dy = dx)                                  */
  double dt;            /* CRM_DT [s]                                            */
  double dt_gl;         /* GCM time step [s]                                     */
  int igstep;           /* get_nstep(): setperturb fires when igstep <= 1        */
  int micro;            /* 0 = sam1mom, 1 = m2005                                */
  int use_crm_accel;    /* mean-state acceleration (accelerate_crm.F90:55)       */
  int crm_accel_uv;
  double crm_accel_factor;
  int twiddle_f32;      /* 1: REAL(4) FFT trig tables like CRM/fft.F90; 0: FP64 */
  int esmt;             /* 1: explicit scalar momentum transport (-DSP_ESMT, scalar_momentum.F90); ABI >= 4 */
} crm_dims_t;

/* crm_input_type (CRM/crm_input_module.F90:10-47) + phys_grid lookups (crm_module.F90:287-288,679) */
typedef struct crm_input_t {
  const double *zmid, *zint;            /* (ld,plev), (ld,plev+1) */
  const double *tl, *ql, *qccl, *qiil;  /* (ld,plev) */
  const double *ps;                     /* (ld) */
  const double *pmid, *pint, *pdel;     /* (ld,plev), (ld,plev+1), (ld,plev) */
  const double *phis;                   /* (ld) */
  const double *ul, *vl;                /* (ld,plev) */
  const double *ocnfrac, *tau00, *wndls, *bflxls;        /* (ld) */
  const double *fluxu00, *fluxv00, *fluxt00, *fluxq00;   /* (ld) */
  const double *lat_deg, *lon_deg;      /* (ld) get_rlat_p/get_rlon_p * 57.296 */
  const int    *gcol;                   /* (ld) get_gcol_p: seed of setperturb */
  const double *ul_esmt, *vl_esmt;      /* (ld,plev) SP_ESMT builds only (crm_input_module.F90:39-42); NULL otherwise (ABI >= 4) */
} crm_input_t;

/* crm_state_type (CRM/crm_state_module.F90:12-56); all (ld,nx,ny,nzm), in/out */
typedef struct crm_state_t {
  double *u_wind, *v_wind, *w_wind, *temperature;
  double *qt, *qp, *qn;                                  /* sam1mom */
  double *nc, *qr, *nr, *qi, *ni, *qs, *ns, *qg, *ng, *qc; /* m2005 (may be NULL for sam1mom) */
} crm_state_t;

/* crm_rad_type (CRM/crm_rad_module.F90:17-34); all (ld,nx_rad,ny_rad,nzm) */
typedef struct crm_rad_t {
  const double *qrad;                         /* in: radiative heating [K/s] */
  double *temperature, *qv, *qc, *qi, *cld;   /* out: time/group means */
  double *nc, *ni, *qs, *ns;                  /* m2005 only */
} crm_rad_t;

/* crm_output_type (CRM/crm_output_module.F90:7-133), same member names */
typedef struct crm_output_t {
  /* (ld,nx,ny,nzm) */
  double *qcl, *qci, *qpl, *qpi, *tk, *tkh;
  /* (ld,nx,ny) */
  double *prec_crm;
  /* (ld) */
  double *cltot, *cllow, *clmed, *clhgh;
  double *precc, *precl, *precsc, *precsl;
  double *jt_crm, *mx_crm;
  double *prectend, *precstend;
  double *taux, *tauy, *z0m, *timing_factor;
  /* (ld,plev) */
  double *cldtop;
  double *qc_mean, *qi_mean, *qs_mean, *qg_mean, *qr_mean;
  double *sltend, *qltend, *qcltend, *qiltend;
  double *cld, *gicewp, *gliqwp;
  double *mctot, *mcup, *mcdn, *mcuup, *mcudn;
  double *mu_crm, *md_crm, *du_crm, *eu_crm, *ed_crm;
  double *flux_qt, *fluxsgs_qt, *tkez, *tkesgsz, *tkz;
  double *flux_u, *flux_v, *flux_qp, *precflux;
  double *qt_ls, *qt_trans, *qp_trans, *qp_fall, *qp_src, *qp_evp, *t_ls;
  /* (ld,plev), reference build option SPMOMTRANS (crm_output_module.F90:76-79, crm_module.F90:1200-1211): resolved
   * convective momentum transport tendencies (uln - ul)/dt_gl, two top CRM levels zeroed; NULL = not wanted (ABI >= 2) */
  double *ultend, *vltend;
  /* ---- m2005 members (crm_output_module.F90:22-31, 60-74; ABI >= 3); ignored for micro = 0, NULL = not wanted ---- */
  /* (ld,nx,ny,nzm): sub-grid vertical velocity sqrt(tke/3) and the process rates of the last sub-step (crm_module.F90:1288-1298) */
  double *wvar, *aut, *acc, *evpc, *evpr, *mlt, *sub, *dep, *con;
  /* (ld,plev): mean number mixing ratios (crm_module.F90:1341-1350, 1379-1384) and time/domain mean process rates (:1386-1414) */
  double *nc_mean, *ni_mean, *ns_mean, *ng_mean, *nr_mean;
  double *aut_a, *acc_a, *evpc_a, *evpr_a, *mlt_a, *sub_a, *dep_a, *con_a;
  /* (ld,plev), SP_ESMT builds (crm_output_module.F90:81-84, crm_module.F90:1185-1192): (mean of the transported scalar momentum -
   * ul_esmt) / dt_gl; required when dims.esmt = 1 (ABI >= 4) */
  double *u_tend_esmt, *v_tend_esmt;
} crm_output_t;

/* Run statistics (new: replaces the GPTL timers 'crm_gpu_region' etc.) */
typedef struct crm_run_stats_t {
  long long nstop;          /* CRM steps taken                              */
  long long ncycle_sum;     /* sum over steps of ncycle = sub-steps taken    */
  long long ncycle_max;
  long long nprec_max;
  long long kernel_launches;/* device kernels launched by this call          */
  double    ms_total;       /* wall/device time of the whole call            */
  double    ms_kernels;     /* device time between first and last kernel     */
  double    ms_h2d, ms_d2h; /* time spent in boundary copies                 */
  long long h2d_bytes, d2h_bytes;
} crm_run_stats_t;

/* ---- entry points --------------------------------------------------------------------- */

/* Create a handle: allocates the persistent workspace arena on `device` for the given
 * geometry (replaces the per-call allocate_*()/prefetch() of crm_module.F90:199-266). */
int crm_b200_init(int device, const crm_dims_t *dims, void **handle);

/* One crm() call (crm_module.F90:45-1718).  `dims` may differ from the init dims only in
 * ncrms (<= init ncrms), ld, dt_gl, igstep and the accel flags. */
int crm_b200_run(void *handle, const crm_dims_t *dims, const crm_input_t *in,
                 crm_state_t *state, crm_rad_t *rad, crm_output_t *out);

int crm_b200_free(void *handle);

/* ---- crm_physics_tend glue around the crm() call (reference PHYS/crm_physics.F90:1006-1011 and 1160-1209) ----------
 * with crm_state / crm_rad / crm_output resident on the GPU across GCM steps these two calls keep
 * the pre/post arithmetic of the caller on the device as well.  Arrays as in crm_b200_run (host or device pointers,
 * leading dimension ld, rows 0..ncrms-1); pdel is state%pdel (ld,plev). */
typedef struct crm_ptend_t {
  double *s;                            /* ptend%s            (ld,plev) <- crm_output%sltend  */
  double *q_vapor, *q_cldliq, *q_cldice;/* ptend%q(:,:,1|ixcldliq|ixcldice) <- qltend, qcltend, qiltend */
  double *prec_dp, *snow_dp;            /* (ld): precc + precl, precsc + precsl (crm_physics.F90:1168-1177) */
  double *cld;                          /* (ld,plev) pbuf CLD <- crm_output%cld (crm_physics.F90:1195)  */
} crm_ptend_t;

/* before crm(): crm_rad%qrad(i,:,:,m) /= pdel(i, plev-m+1), m = 1..crm_nz  ("for energy conservation", :1006-1011) */
int crm_b200_physics_pre(void *handle, const crm_dims_t *dims, const double *pdel, double *qrad);

/* after crm(): ptend assembly, the convective/stratiform precipitation merge (crm_output%precc += precl; precl = 0; the same
 * for precsc/precsl, updated in place in `out`), CLD, and qrad *= pdel (:1160-1209).  NULL members of ptend are skipped. */
int crm_b200_physics_post(void *handle, const crm_dims_t *dims, const double *pdel, double *qrad, crm_output_t *out,
                          crm_ptend_t *ptend);

/* ---- device-resident crm_state / crm_rad across GCM steps, keyed by the physics chunk (ABI >= 5) -----------------------
 * SURVEY.md 8(f) items 1 and 3.  In the reference crm_state%* and crm_rad%* are pointers into the physics buffer of chunk
 * `lchnk` (PHYS/crm_physics.F90:126-147 pbuf_add_field, :811-828 pbuf_get_field): they persist from one GCM step to the
 * next and nothing but crm() (and the radiation, for crm_rad) touches them, yet every call moves them across the host
 * boundary (7 x 1.44 GB each way at BASELINE config 2).  Here the handle keeps one copy per chunk in HBM, in the boundary
 * layout with the leading dimension given at `put`:
 *   put   after the caller has seeded the state on the first step (crm_physics.F90:912-939) or read a restart file;
 *         NULL members are zero on creation and left alone on later puts;
 *   run   one crm() call on the resident state: only crm_input, (optionally) qrad and the outputs cross the boundary.
 *         `rad` may be NULL; a non-NULL rad->qrad replaces the resident qrad before the call, non-NULL output members of
 *         `rad` receive copies of the resident results after it.  dims->ncrms and dims->ld must equal those of `put`;
 *   get   copies resident members into the caller's arrays (restart write; NULL members skipped; either struct may be NULL);
 *   field device pointer of one resident member for callers that run their own device code on it (a GPU radiation
 *         reading crm_rad%*, crm_b200_physics_pre/post on "qrad"): crm_state member names, crm_rad members prefixed
 *         "rad_" ("rad_temperature", ...) and "qrad"; NULL if unknown;
 *   drop  frees one chunk, or all of them with lchnk < 0 (crm_b200_free drops them too).
 * Several chunks may share one handle (one workspace arena): calls on a handle are serialised anyway. */
int crm_b200_resident_put(void *handle, int lchnk, const crm_dims_t *dims, const crm_state_t *state, const crm_rad_t *rad);
int crm_b200_resident_run(void *handle, int lchnk, const crm_dims_t *dims, const crm_input_t *in, crm_rad_t *rad,
                          crm_output_t *out);
int crm_b200_resident_get(void *handle, int lchnk, const crm_dims_t *dims, crm_state_t *state, crm_rad_t *rad);
double *crm_b200_resident_field(void *handle, int lchnk, const char *name);
int crm_b200_resident_drop(void *handle, int lchnk);
long long crm_b200_resident_bytes(void *handle);   /* HBM held by resident chunks of this handle */

/* CRM-level history stream of crm_physics_tend (PHYS/crm_physics.F90:1214-1273, the `outfld('CRM_*')` calls) for a resident
 * chunk: writes the named field in the boundary layout (dims->ld, nx, ny, nzm) into `buf` (host or device).  State fields
 * "CRM_U","CRM_V","CRM_W","CRM_T" (+ m2005 "CRM_NC","CRM_NI","CRM_NR","CRM_NS","CRM_NG","CRM_QR","CRM_QS","CRM_QG") come
 * from the resident copy; "CRM_QV" (= qt - qcl - qci, m2005: qt - qcl, formed on the device), "CRM_QC","CRM_QI","CRM_QPC",
 * "CRM_QPI","CRM_TK","CRM_TKH" (+ m2005 "CRM_AUT","CRM_ACC","CRM_EVPC","CRM_EVPR","CRM_MLT","CRM_SUB","CRM_DEP","CRM_CON")
 * come from the workspace of the last crm_b200_resident_run and are refused (rc 54) if that run was not on `lchnk`. */
int crm_b200_history_field(void *handle, int lchnk, const crm_dims_t *dims, const char *name, double *buf);

/* Message of the last failure on this handle (or of a failed init if handle==NULL). */
const char *crm_b200_last_error(void *handle);

/* Statistics of the most recent crm_b200_run on this handle. */
int crm_b200_get_stats(void *handle, crm_run_stats_t *stats);

/* Test/diagnostic hook: copy an internal field, as left by the last run, into a host buffer
 * in the boundary layout (ncrms,nx,ny,nz[m]) with ld = ncrms.  Names: "u","w","t","qt","qp",
 * "qn","tabs","qv","qcl","qci","qpl","qpi","tk","tkh","tke", the Adams-Bashforth tendency slots
 * "dudt","dwdt" (slot just computed), "dudt_a","dwdt_a" (na), "dudt_c","dwdt_c" (nc), the mass
 * fluxes "umf","wmf", and for 3-D CRMs "v","dvdt","dvdt_a","dvdt_c","vmf".  (The pressure is not
 * kept on the device: it lives in shared memory between the forward and backward solve kernels.)
 * With the environment variable CRM_DEBUG_STOP=1..8 crm_b200_run stops inside the first sub-step
 * after the phase of that number (forcing, sgs, momentum, pressure+adams, scalar advection,
 * precipitation fall+uvw, scalar diffusion, microphysics) so the per-phase parity tests can fetch
 * the intermediate state.  Returns the number of doubles written, or <0 on error
 * (-3 unknown name, -2 buffer too small). */
long long crm_b200_debug_fetch(void *handle, const char *name, double *host_buf, size_t n);

/* Measurement hook: when on, every kernel launch of crm_b200_run is bracketed by CUDA events (this serialises the
 * launch queue) and crm_b200_kernel_times returns "kernel_name total_ms launches\n" lines for the last run. */
int crm_b200_set_kernel_timing(void *handle, int on);
const char *crm_b200_kernel_times(void *handle);

int crm_b200_abi_version(void);

/* sizeof() of the boundary structs as compiled into the library, for binding-side layout checks:
 * which = 0 crm_dims_t, 1 crm_input_t, 2 crm_state_t, 3 crm_rad_t, 4 crm_output_t, 5 crm_run_stats_t; -1 otherwise. */
long long crm_b200_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* CRM_B200_H */
