/*
 * crm_replay.c -- plain-C host of the C ABI (include/crm_b200.h): replays one recorded crm() call.
 *
 * The reference's hot path is entered from compiled Fortran (PHYS/crm_physics.F90:1145-1156); this program is the same call made
 * from a compiled host language that IS available in the development image (C, gcc): no Python, no torch, nothing but the header.
 * It reads a replay file holding the arguments of one crm() call (every member array of crm_input / crm_state / crm_rad and the
 * shapes of crm_output), calls crm_b200_init / crm_b200_run (ncalls times, chaining crm_state like consecutive GCM steps) /
 * crm_b200_free, and writes crm_state, crm_rad and crm_output back in the same format.  BASELINE config 1 is specified as
 * "crm() I/O replayed": acme-ecp_b200/replay.py writes and reads the files, tests/test_replay_c.py compares the result with the
 * ctypes path bit for bit.
 *
 * The library is loaded with dlopen (first argument), so the one binary runs against acme-ecp_b200/csrc/libcrm_b200.so (GPU) and
 * against the dev-only host emulation build of the same sources (tests/emu), which is how the CPU-only test suite exercises it.
 *
 *   gcc -O2 -std=c11 -I include examples/crm_replay.c -o examples/crm_replay -ldl
 *   examples/crm_replay <lib.so> <in.crmr> <out.crmr> [ncalls]
 *
 * File format (little endian): "CRMRPLY1", int32 sizeof(crm_dims_t), crm_dims_t, int32 nrec, then nrec records of
 *   char name[40] ("input.tl", "state.qt", "rad.qrad", "output.sltend", ...), int32 dtype (0 = f64, 1 = i32), int32 has_data,
 *   int64 count, [count elements if has_data].
 */
#include <dlfcn.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crm_b200.h"

/* member tables: name -> offset of the pointer inside the boundary struct (order = header order; tests/test_replay_c.py checks them
 * against include/crm_b200.h) */
#define M(s, m) {#m, offsetof(s, m)}
typedef struct { const char *name; size_t off; } member_t;
static const member_t k_input[] = {
    M(crm_input_t, zmid), M(crm_input_t, zint), M(crm_input_t, tl), M(crm_input_t, ql), M(crm_input_t, qccl), M(crm_input_t, qiil),
    M(crm_input_t, ps), M(crm_input_t, pmid), M(crm_input_t, pint), M(crm_input_t, pdel), M(crm_input_t, phis), M(crm_input_t, ul),
    M(crm_input_t, vl), M(crm_input_t, ocnfrac), M(crm_input_t, tau00), M(crm_input_t, wndls), M(crm_input_t, bflxls),
    M(crm_input_t, fluxu00), M(crm_input_t, fluxv00), M(crm_input_t, fluxt00), M(crm_input_t, fluxq00), M(crm_input_t, lat_deg),
    M(crm_input_t, lon_deg), M(crm_input_t, gcol), M(crm_input_t, ul_esmt), M(crm_input_t, vl_esmt)};
static const member_t k_state[] = {
    M(crm_state_t, u_wind), M(crm_state_t, v_wind), M(crm_state_t, w_wind), M(crm_state_t, temperature), M(crm_state_t, qt),
    M(crm_state_t, qp), M(crm_state_t, qn), M(crm_state_t, nc), M(crm_state_t, qr), M(crm_state_t, nr), M(crm_state_t, qi),
    M(crm_state_t, ni), M(crm_state_t, qs), M(crm_state_t, ns), M(crm_state_t, qg), M(crm_state_t, ng), M(crm_state_t, qc)};
static const member_t k_rad[] = {
    M(crm_rad_t, qrad), M(crm_rad_t, temperature), M(crm_rad_t, qv), M(crm_rad_t, qc), M(crm_rad_t, qi), M(crm_rad_t, cld),
    M(crm_rad_t, nc), M(crm_rad_t, ni), M(crm_rad_t, qs), M(crm_rad_t, ns)};
static const member_t k_output[] = {
    M(crm_output_t, qcl), M(crm_output_t, qci), M(crm_output_t, qpl), M(crm_output_t, qpi), M(crm_output_t, tk), M(crm_output_t, tkh),
    M(crm_output_t, prec_crm), M(crm_output_t, cltot), M(crm_output_t, cllow), M(crm_output_t, clmed), M(crm_output_t, clhgh),
    M(crm_output_t, precc), M(crm_output_t, precl), M(crm_output_t, precsc), M(crm_output_t, precsl), M(crm_output_t, jt_crm),
    M(crm_output_t, mx_crm), M(crm_output_t, prectend), M(crm_output_t, precstend), M(crm_output_t, taux), M(crm_output_t, tauy),
    M(crm_output_t, z0m), M(crm_output_t, timing_factor), M(crm_output_t, cldtop), M(crm_output_t, qc_mean), M(crm_output_t, qi_mean),
    M(crm_output_t, qs_mean), M(crm_output_t, qg_mean), M(crm_output_t, qr_mean), M(crm_output_t, sltend), M(crm_output_t, qltend),
    M(crm_output_t, qcltend), M(crm_output_t, qiltend), M(crm_output_t, cld), M(crm_output_t, gicewp), M(crm_output_t, gliqwp),
    M(crm_output_t, mctot), M(crm_output_t, mcup), M(crm_output_t, mcdn), M(crm_output_t, mcuup), M(crm_output_t, mcudn),
    M(crm_output_t, mu_crm), M(crm_output_t, md_crm), M(crm_output_t, du_crm), M(crm_output_t, eu_crm), M(crm_output_t, ed_crm),
    M(crm_output_t, flux_qt), M(crm_output_t, fluxsgs_qt), M(crm_output_t, tkez), M(crm_output_t, tkesgsz), M(crm_output_t, tkz),
    M(crm_output_t, flux_u), M(crm_output_t, flux_v), M(crm_output_t, flux_qp), M(crm_output_t, precflux), M(crm_output_t, qt_ls),
    M(crm_output_t, qt_trans), M(crm_output_t, qp_trans), M(crm_output_t, qp_fall), M(crm_output_t, qp_src), M(crm_output_t, qp_evp),
    M(crm_output_t, t_ls), M(crm_output_t, ultend), M(crm_output_t, vltend), M(crm_output_t, wvar), M(crm_output_t, aut),
    M(crm_output_t, acc), M(crm_output_t, evpc), M(crm_output_t, evpr), M(crm_output_t, mlt), M(crm_output_t, sub), M(crm_output_t, dep),
    M(crm_output_t, con), M(crm_output_t, nc_mean), M(crm_output_t, ni_mean), M(crm_output_t, ns_mean), M(crm_output_t, ng_mean),
    M(crm_output_t, nr_mean), M(crm_output_t, aut_a), M(crm_output_t, acc_a), M(crm_output_t, evpc_a), M(crm_output_t, evpr_a),
    M(crm_output_t, mlt_a), M(crm_output_t, sub_a), M(crm_output_t, dep_a), M(crm_output_t, con_a), M(crm_output_t, u_tend_esmt),
    M(crm_output_t, v_tend_esmt)};
#define COUNT(a) (sizeof(a) / sizeof((a)[0]))

typedef struct { char name[40]; int32_t dtype, has_data; int64_t count; void *data; } record_t;

static int bind(void *strct, const member_t *tab, size_t n, const char *member, void *data) {
  for (size_t q = 0; q < n; q++)
    if (!strcmp(tab[q].name, member)) { *(void **)((char *)strct + tab[q].off) = data; return 0; }
  return 1;
}

#define DIE(...) do { fprintf(stderr, "crm_replay: " __VA_ARGS__); fprintf(stderr, "\n"); return 2; } while (0)

int main(int argc, char **argv) {
  if (argc < 4) DIE("usage: crm_replay <libcrm_b200.so> <in.crmr> <out.crmr> [ncalls]");
  const int ncalls = argc > 4 ? atoi(argv[4]) : 1;
  void *lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) DIE("cannot load %s: %s", argv[1], dlerror());
  int (*p_init)(int, const crm_dims_t *, void **) = (int (*)(int, const crm_dims_t *, void **))dlsym(lib, "crm_b200_init");
  int (*p_run)(void *, const crm_dims_t *, const crm_input_t *, crm_state_t *, crm_rad_t *, crm_output_t *) =
      (int (*)(void *, const crm_dims_t *, const crm_input_t *, crm_state_t *, crm_rad_t *, crm_output_t *))dlsym(lib, "crm_b200_run");
  int (*p_free)(void *) = (int (*)(void *))dlsym(lib, "crm_b200_free");
  const char *(*p_err)(void *) = (const char *(*)(void *))dlsym(lib, "crm_b200_last_error");
  int (*p_abi)(void) = (int (*)(void))dlsym(lib, "crm_b200_abi_version");
  long long (*p_sizeof)(int) = (long long (*)(int))dlsym(lib, "crm_b200_sizeof");
  int (*p_stats)(void *, crm_run_stats_t *) = (int (*)(void *, crm_run_stats_t *))dlsym(lib, "crm_b200_get_stats");
  if (!p_init || !p_run || !p_free || !p_err || !p_abi || !p_sizeof || !p_stats) DIE("%s does not export the C ABI of crm_b200.h", argv[1]);
  /* binding-side layout checks (what the Fortran shim asserts as well) */
  if (p_abi() != CRM_B200_ABI_VERSION) DIE("ABI version %d, header says %d", p_abi(), CRM_B200_ABI_VERSION);
  if (p_sizeof(0) != (long long)sizeof(crm_dims_t) || p_sizeof(1) != (long long)sizeof(crm_input_t) || p_sizeof(2) != (long long)sizeof(crm_state_t) ||
      p_sizeof(3) != (long long)sizeof(crm_rad_t) || p_sizeof(4) != (long long)sizeof(crm_output_t) || p_sizeof(5) != (long long)sizeof(crm_run_stats_t))
    DIE("struct layout of the library differs from the header this program was compiled with");

  FILE *fi = fopen(argv[2], "rb");
  if (!fi) DIE("cannot open %s", argv[2]);
  char magic[8];
  int32_t sz = 0, nrec = 0;
  crm_dims_t dims;
  if (fread(magic, 1, 8, fi) != 8 || memcmp(magic, "CRMRPLY1", 8)) DIE("%s is not a replay file", argv[2]);
  if (fread(&sz, 4, 1, fi) != 1 || sz != (int32_t)sizeof(crm_dims_t)) DIE("replay file was written for another crm_dims_t");
  if (fread(&dims, sizeof(dims), 1, fi) != 1 || fread(&nrec, 4, 1, fi) != 1 || nrec < 1 || nrec > 4096) DIE("truncated replay file");
  record_t *rec = (record_t *)calloc((size_t)nrec, sizeof(record_t));
  crm_input_t in; crm_state_t st; crm_rad_t rad; crm_output_t out;
  memset(&in, 0, sizeof(in)); memset(&st, 0, sizeof(st)); memset(&rad, 0, sizeof(rad)); memset(&out, 0, sizeof(out));
  for (int r = 0; r < nrec; r++) {
    record_t *c = &rec[r];
    if (fread(c->name, 1, 40, fi) != 40 || fread(&c->dtype, 4, 1, fi) != 1 || fread(&c->has_data, 4, 1, fi) != 1 || fread(&c->count, 8, 1, fi) != 1)
      DIE("truncated record %d", r);
    c->name[39] = 0;
    const size_t esz = c->dtype == 1 ? 4 : 8;
    c->data = calloc((size_t)c->count ? (size_t)c->count : 1, esz);
    if (!c->data) DIE("out of memory");
    if (c->has_data && fread(c->data, esz, (size_t)c->count, fi) != (size_t)c->count) DIE("truncated data of %s", c->name);
    const char *dot = strchr(c->name, '.');
    if (!dot) DIE("bad record name %s", c->name);
    int bad;
    if (!strncmp(c->name, "input.", 6)) bad = bind(&in, k_input, COUNT(k_input), dot + 1, c->data);
    else if (!strncmp(c->name, "state.", 6)) bad = bind(&st, k_state, COUNT(k_state), dot + 1, c->data);
    else if (!strncmp(c->name, "rad.", 4)) bad = bind(&rad, k_rad, COUNT(k_rad), dot + 1, c->data);
    else if (!strncmp(c->name, "output.", 7)) bad = bind(&out, k_output, COUNT(k_output), dot + 1, c->data);
    else bad = 1;
    if (bad) DIE("record %s names no member of the boundary structs", c->name);
  }
  fclose(fi);

  void *h = NULL;
  int rc = p_init(0, &dims, &h);
  if (rc) DIE("crm_b200_init failed (%d): %s", rc, p_err(NULL));
  crm_run_stats_t stats;
  memset(&stats, 0, sizeof(stats));
  for (int call = 0; call < ncalls; call++) {
    crm_dims_t d = dims;
    if (call > 0 && d.igstep <= 1) d.igstep = 2;          /* setperturb fires on the first GCM step only */
    rc = p_run(h, &d, &in, &st, &rad, &out);
    if (rc) { fprintf(stderr, "crm_replay: crm_b200_run failed (%d): %s\n", rc, p_err(h)); p_free(h); return 3; }
    p_stats(h, &stats);
    printf("call %d: nstop %lld  sub-steps %lld  kernel launches %lld  %.3f ms\n", call, stats.nstop, stats.ncycle_sum, stats.kernel_launches,
           stats.ms_total);
  }
  p_free(h);

  FILE *fo = fopen(argv[3], "wb");
  if (!fo) DIE("cannot open %s", argv[3]);
  int32_t nout = 0;
  for (int r = 0; r < nrec; r++) if (strncmp(rec[r].name, "input.", 6)) nout++;
  fwrite("CRMRPLY1", 1, 8, fo); fwrite(&sz, 4, 1, fo); fwrite(&dims, sizeof(dims), 1, fo); fwrite(&nout, 4, 1, fo);
  for (int r = 0; r < nrec; r++) {
    record_t *c = &rec[r];
    if (!strncmp(c->name, "input.", 6)) continue;
    int32_t one = 1;
    fwrite(c->name, 1, 40, fo); fwrite(&c->dtype, 4, 1, fo); fwrite(&one, 4, 1, fo); fwrite(&c->count, 8, 1, fo);
    fwrite(c->data, c->dtype == 1 ? 4 : 8, (size_t)c->count, fo);
  }
  fclose(fo);
  for (int r = 0; r < nrec; r++) free(rec[r].data);
  free(rec);
  dlclose(lib);
  return 0;
}
