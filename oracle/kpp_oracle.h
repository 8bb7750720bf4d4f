/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 * A plain-C restatement of the reference's KPP Rosenbrock chemistry path
 * (geoschem/geos-chem, KPP/<mech>/gckpp_*.F90).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (geos_chem_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED for the fullchem integrator by the reference's own known-answer
 * vector KPP/standalone/Beijing_L1_20190701_0040.txt (12 internal steps, Hexit 497.8023 s,
 * A(1:1058) bit-equal) -- see tests/test_oracle_golden.py.  UNPINNED by the reference for:
 * gas-phase rate laws beyond ~1e-5 (fixture prints T to 2 decimals), Hg, carbon, auto-reduce.
 */
#ifndef KPP_ORACLE_H
#define KPP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* per-cell meteorological scalars of commonIncludeVars.H as set by Set_Kpp_GridBox_Values
 * (GeosCore/fullchem_mod.F90:2089-2167) */
typedef struct {
  double TEMP, NUMDEN, H2O, PRESS;
  double INV_TEMP, TEMP_OVER_K300, K300_OVER_TEMP, SR_TEMP;
} kpp_met_t;

typedef struct kpp_mech {
  const char *name;
  int nvar, nfix, nspec, nreact, lu_nonzero, nb, nphot, next;
  const int *lu_crow, *lu_diag, *lu_icol;   /* 0-based CSR of gckpp_JacobianSP.F90 */
  int fun_is_split;  /* FunTemplate calls Fun_SPLIT (fullchem) or the aggregate Fun (Hg, carbon) */
  void (*rates)(const double *V, const double *F, const double *RCT, double *A);
  void (*fun_split)(const double *V, const double *F, const double *RCT, double *Vdot,
                    double *P_VAR, double *D_VAR, double *A, const unsigned char *do_fun);
  void (*fun)(const double *V, const double *F, const double *RCT, double *Vdot, double *A);
  void (*jac_sp)(const double *V, const double *F, const double *RCT, double *JVS, double *B,
                 const unsigned char *do_jvs);
  void (*solve)(const double *JVS, double *X, const unsigned char *do_slv);
  void (*update_rconst)(const kpp_met_t *m, const double *PHOTOL, const double *khet, double *RCONST);
} kpp_mech_t;

extern const kpp_mech_t kpp_mech_fullchem, kpp_mech_Hg, kpp_mech_carbon;

/* mech_id: 0 fullchem, 1 Hg, 2 carbon */
const kpp_mech_t *kpp_oracle_mech(int mech_id);
int kpp_oracle_dims(int mech_id, int *dims /* nvar,nfix,nspec,nreact,lu_nonzero,nphot,next */);

/* Integrate (gckpp_Integrator.F90:80-162) for one cell, AoS arrays like the Fortran globals.
 * C[nspec] in/out, RCONST[nreact], ATOL/RTOL[nvar], ICNTRL_U[20], RCNTRL_U[20] ->
 * ISTATUS[20], RSTATUS[20]; returns IERR. */
int kpp_oracle_integrate_cell(int mech_id, double tin, double tout, double *C, const double *RCONST,
                              const double *ATOL, const double *RTOL, const int *ICNTRL_U,
                              const double *RCNTRL_U, int *ISTATUS, double *RSTATUS);

/* Batched driver with the GPU ABI's cell-fastest layout and Do_FullChem's OpenMP loop
 * (GeosCore/fullchem_mod.F90:528-546: schedule(dynamic,24)).  hstart may be NULL.
 * istatus [8][ncell], rstatus [4][ncell], ierr [ncell]. nthreads<=0 -> omp default. */
int kpp_oracle_integrate(int mech_id, int ncell, double tin, double tout, const double *conc_in,
                         const double *rconst, const double *atol, const double *rtol,
                         const int *icntrl, const double *rcntrl, const double *hstart,
                         double *conc_out, int *istatus, double *rstatus, int *ierr, int nthreads);

/* Update_RCONST (gckpp_Rates.F90:408-1503) over cells, cell-fastest arrays:
 * temp/numden/h2o [ncell], photol [nphot][ncell], khet [next][ncell] -> rconst [nreact][ncell] */
int kpp_oracle_update_rconst(int mech_id, int ncell, const double *temp, const double *numden,
                             const double *h2o, const double *photol, const double *khet,
                             double *rconst, int nthreads);

/* keepActive / keepSpcActive of the auto-reduce solver (fullchem_AutoReduceFuncs.F90:40-140): 0-based species
 * indices that are never removed; n = 0 switches keepActive off.  Global (like the Fortran module variable). */
int kpp_oracle_set_keep_active(int mech_id, int n, const int *idx0);

/* single-cell pieces for unit tests */
int kpp_oracle_fun(int mech_id, const double *C, const double *RCONST, double *Vdot, double *A);
int kpp_oracle_jac(int mech_id, const double *C, const double *RCONST, double *JVS);
int kpp_oracle_decomp(int mech_id, double *JVS);                     /* returns IER (0 ok) */
int kpp_oracle_solve(int mech_id, const double *JVS, double *X);

#ifdef __cplusplus
}
#endif
#endif
