// Mechanism tables shared by the host (static data emitted by kppgen/emit_cuda.py) and the
// table-driven kernels.  Term encoding: see emit_cuda.py.
#pragma once
#include <stdint.h>

struct gckpp_host_tables_t {
  const char *name;
  int nvar, nfix, nspec, nreact, nnz, nb, nphot, next, fun_split;
  const int *a_term;                                   // [nreact][4]
  const int *p_ptr; const double *p_coef; const int *p_rxn; int np;   // production  (Fun_SPLIT P_VAR)
  const int *d_ptr; const int *d_term; int nd;         // destruction (Fun_SPLIT D_VAR) [nd][4]
  const int *v_ptr; const double *v_coef; const int *v_rxn; int nv;   // aggregate Fun Vdot
  const int *crow, *diag, *icol;                       // CSR of LU_CROW/LU_DIAG/LU_ICOL (0-based)
  const int *b_term;                                   // [nb][4]   Jac_SP partials
  const int *j_ptr; const double *j_coef; const int *j_b; int nj;     // JVS sums
  const double *lit; int nlit;
  const double *ohr_coef; const int *ohr_rxn, *ohr_spc; int nohr;   // Get_OHreactivity: sum coef * RCONST(rxn) [* C(spc)]
};

// Round/bundle schedules of the shared-memory kernel (kppgen/sched.py documents the encoding).
struct gckpp_sched_tables_t {
  int nrows, nbundles, nrounds;
  const uint32_t *chunks /* [nrows][32][4] */, *brow /* [nbundles+1] first chunk row of a bundle */,
      *rounds /* [nrounds][3]: first bundle, last+1, kind (bit 4: LU division round) */;
  int phase[12];                                       // round ranges of vdot, jvs, lu, scale, fwd, bwd
  const double *coefs; int ncoef;
  const uint16_t *tpos;                                // [32][32] tposT[j][i] = position of G(h+i,h+j) or 0xFFFF
  int head, tail;
};

// Device copy: same fields, device pointers.
struct MechDev {
  int nvar, nfix, nspec, nreact, nnz, nb, nphot, next, fun_split;
  const int4 *a_term;
  const int *p_ptr; const double *p_coef; const int *p_rxn;
  const int *d_ptr; const int4 *d_term;
  const int *v_ptr; const double *v_coef; const int *v_rxn;
  const int *crow, *diag, *icol;
  const int4 *b_term;
  const int *j_ptr; const double *j_coef; const int *j_b;
  const double *lit;
};

// Per-cell meteorological scalars (commonIncludeVars.H) as Set_Kpp_GridBox_Values derives them
// (GeosCore/fullchem_mod.F90:2139-2150).
struct MetCell {
  double TEMP, NUMDEN, H2O;
  double INV_TEMP, TEMP_OVER_K300, K300_OVER_TEMP, SR_TEMP;
  double FOUR_R_T, FOUR_RGASLATM_T, EIGHT_RSTARG_T, RELHUM;     // :2145-2160, read by the heterogeneous laws
};
