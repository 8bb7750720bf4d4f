/* gckpp_gpu.h -- C ABI of the B200 KPP chemistry path (libgckpp_b200.so).
 *
 * Drop-in, batched replacement for what GEOS-Chem's chemistry drivers call per grid box:
 *
 *   Update_RCONST()                         KPP/<mech>/gckpp_Rates.F90:408
 *   Integrate(TIN,TOUT,ICNTRL_U,RCNTRL_U,   KPP/<mech>/gckpp_Integrator.F90:80-81
 *             ISTATUS_U,RSTATUS_U,IERR_U)
 *   Fun(V,F,RCT,Vdot,Aout)                  KPP/<mech>/gckpp_Function.F90:51
 *
 * called from Do_FullChem (GeosCore/fullchem_mod.F90:953,1033,1157,1162), ChemMercury
 * (GeosCore/mercury_mod.F90:1106,1132) and Chem_Carbon_Gases (GeosCore/carbon_gases_mod.F90:532,539).
 * The reference passes state through OpenMP-threadprivate module variables (C, RCONST, TEMP,
 * NUMDEN, H2O, PHOTOL, K_MT, K_CLD, ATOL, RTOL: gckpp_Global.F90:47-80); here the same
 * quantities are arrays over cells.  One call covers a whole chemistry step on one GPU.
 *
 * Conventions
 *   - ISO_C_BINDING compatible: scalars by value, arrays as raw pointers, no descriptors.
 *   - Every per-cell array is CELL-FASTEST: element (k, cell) at  a[k*ncell + cell] .  This is
 *     the memory order of the 356 separate State_Chm%Species(n)%Conc(I,J,L) arrays with
 *     cell = I + NX*(J-1) + NX*NY*(L-1)  (0-based here).
 *   - Species order = the mechanism's SPC_NAMES (variable species first, then fixed).
 *   - The caller owns every array passed in; nothing is retained after return.
 *   - Return value: 0 ok; >0 = number of cells that failed twice (retry enabled);
 *     <0 = -(1000 + cudaError) for CUDA errors, -1..-5 option errors (same codes as the
 *     reference's IERR), -10 bad argument, -11 library built without the requested mechanism.
 *   - Per-cell ierr[] uses the reference's codes (gckpp_Integrator.F90:551-571): 1 success,
 *     -6 too many steps, -7 step too small, -8 matrix repeatedly singular; 0 = cell not in
 *     the chemistry grid (active[cell]==0).
 *   - ICNTRL/RCNTRL have the reference's meaning (gckpp_Integrator.F90:165-260): zero /
 *     non-positive entries mean "default" (Integrate's WHERE merge, :112-117).  ICNTRL(15)
 *     must be -1 (rates are never refreshed inside the integrator, as in GEOS-Chem).
 */
#ifndef GCKPP_GPU_H
#define GCKPP_GPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GCKPP_MECH_FULLCHEM 0
#define GCKPP_MECH_HG       1
#define GCKPP_MECH_CARBON   2

typedef struct gckpp_gpu_handle gckpp_gpu_handle_t;

/* dims[0..6] = NVAR, NFIX, NSPEC, NREACT, LU_NONZERO, NPHOT (PHOTOL entries read), NEXT
 * (rate constants supplied by the caller: K_MT, K_CLD and the State_Het-dependent laws, in
 * ascending reaction order; fullchem: 113).  (gckpp_Parameters.F90:34-54) */
int gckpp_gpu_dims(int mech_id, int32_t *dims);

/* Species name of index i (0-based; gckpp_Monitor.F90 SPC_NAMES); NULL if out of range. */
const char *gckpp_gpu_spc_name(int mech_id, int i);

/* Create a solver instance on CUDA device `device`, sized for up to max_cells per call
 * (a sizing hint: buffers grow on demand).  Replaces nothing in the reference (KPP has no
 * state beyond its module variables); owns device buffers and a stream. */
int gckpp_gpu_init(int mech_id, int device, int max_cells, gckpp_gpu_handle_t **handle);
int gckpp_gpu_finalize(gckpp_gpu_handle_t *handle);

/* Options (integer-valued; an unknown key returns -10):
 *   "retry"      1 = on IERR<0 redo the cell once with Hstart=0 from the saved
 *                concentrations (Do_FullChem's policy, fullchem_mod.F90:1138-1162); default 0
 *   "kernel"     -1 = choose (default): the shared-memory block kernel for fullchem with ICNTRL(3) = 0 or 4 (Rodas3), with or
 *                without auto-reduce; the unrolled one-cell-per-thread kernel for Hg; the table-driven kernel otherwise;
 *                0 = table-driven one-cell-per-lane kernel that keeps the reference's operation order
 *                (bit-identical step sequences with the CPU restatement, any ICNTRL(3) method, auto-reduce);
 *                1 = shared-memory-resident Rodas3 block kernel (sums re-associated, FMA contraction, rounding-level
 *                differences); 2 = warp-group kernel (results vary run to run at 1e-9: on request only);
 *                3 = lane kernel (one cell per lane, streamed workspace, every method); 4 = unrolled kernel (Hg)
 *   "wave_cells"        host-buffer entry: cells per wave (0 = choose from the free device memory); the copies of a wave
 *                       overlap the integration of its neighbours
 *   "device_wave_cells" device entry without rate constants: cells per Update_RCONST wave (bounds the RCONST scratch)
 *   "pin"        1 = page-lock the caller's pageable host arrays for the duration of the call
 *   "blocks_per_sm", "blocks_cap", "threads", "chunks"   launch-shape knobs of the table-driven kernel and test hooks
 */
int gckpp_gpu_set_option(gckpp_gpu_handle_t *handle, const char *key, int value);

/* Update_RCONST + Integrate for ncell cells.  HOST pointers; copies in/out internally.
 *   conc_in   [NSPEC][ncell]  concentrations, molec/cm3            (C of gckpp_Global)
 *   rconst    [NREACT][ncell] rate constants, or NULL to compute them on the device from
 *             temp/numden/h2o/photol/khet (Update_RCONST)
 *   temp,numden,h2o [ncell]   TEMP [K], NUMDEN [molec/cm3], H2O [molec/cm3]
 *                             (Set_Kpp_GridBox_Values, fullchem_mod.F90:2139-2141)
 *   photol    [NPHOT][ncell]  PHOTOL(1:NPHOT) J-values [1/s]  (fullchem_mod.F90:621-642)
 *   khet      [NEXT][ncell]   externally supplied rate constants (see gckpp_gpu_dims)
 *   atol,rtol [NVAR]          shared tolerances (gckpp_Global.F90:78-80)
 *   icntrl[20], rcntrl[20]    integrator options
 *   hstart    [ncell] or NULL per-cell RCNTRL(3) = State_Chm%KPPHvalue (fullchem_AutoReduceFuncs.F90:287)
 *   active    [ncell] or NULL 0 = skip cell (not InChemGrid, fullchem_mod.F90:804)
 *   conc_out  [NSPEC][ncell]  raw integrator output (no MAX(C,0) clip: fullchem_mod.F90:1345 is the caller's)
 *   istatus   [8][ncell]      Nfun,Njac,Nstp,Nacc,Nrej,Ndec,Nsol,Nsng   (may be NULL)
 *   rstatus   [4][ncell]      Texit,Hexit,Hnew,ARthr                    (may be NULL)
 *   ierr      [ncell]                                                    (may be NULL)
 */
int gckpp_gpu_integrate(gckpp_gpu_handle_t *handle, int ncell, double tin, double tout,
                        const double *conc_in, const double *rconst,
                        const double *temp, const double *numden, const double *h2o,
                        const double *photol, const double *khet,
                        const double *atol, const double *rtol,
                        const int32_t *icntrl, const double *rcntrl,
                        const double *hstart, const uint8_t *active,
                        double *conc_out, int32_t *istatus, double *rstatus, int32_t *ierr);

/* Same, with every per-cell array already resident in device memory (atol/rtol/icntrl/rcntrl
 * stay host pointers).  Asynchronous work is ordered on the handle's stream and the call
 * returns after the stream is synchronised.  rconst_work: optional device scratch
 * [NREACT][ncell] used when rconst == NULL (NULL = library allocates). */
int gckpp_gpu_integrate_device(gckpp_gpu_handle_t *handle, int ncell, double tin, double tout,
                               const double *conc_in, const double *rconst,
                               const double *temp, const double *numden, const double *h2o,
                               const double *photol, const double *khet,
                               const double *atol, const double *rtol,
                               const int32_t *icntrl, const double *rcntrl,
                               const double *hstart, const uint8_t *active,
                               double *conc_out, int32_t *istatus, double *rstatus, int32_t *ierr);

/* Update_RCONST alone (RxnConst diagnostic, fullchem_mod.F90:997-1002): rconst_out [NREACT][ncell]. */
int gckpp_gpu_update_rconst(gckpp_gpu_handle_t *handle, int ncell,
                            const double *temp, const double *numden, const double *h2o,
                            const double *photol, const double *khet, double *rconst_out);
int gckpp_gpu_update_rconst_device(gckpp_gpu_handle_t *handle, int ncell,
                                   const double *temp, const double *numden, const double *h2o,
                                   const double *photol, const double *khet, double *rconst_out);

/* Heterogeneous rate laws on the device (fullchem; SURVEY 8 f1, first part).  Without these calls every externally
 * supplied constant comes from khet, as before.  With them, Update_RCONST -- inside gckpp_gpu_integrate[_device] and in
 * the stand-alone entry points -- evaluates 61 of the 113 itself (KPP/fullchem/fullchem_RateLawFuncs.F90: VOCuptk1stOrd,
 * IEPOXuptk1stOrd, MGLYuptk1stOrd, GLYXuptk1stOrd :3286-3458; Iuptk* / IbrkdnByAcid* :2371-2542; HO2uptk1stOrd :1461-1484;
 * HBrUptkBySALA/SALC :1423-1455; OHuptkBySALACl/SALCCl :3244-3280; utilities rateLawUtilFuncs.F90:77-140, 459-495) and
 * takes only the remaining ones (K_MT, K_CLD, cloud / halogen / N2O5 / NO2 / NO3 laws) from khet.
 *   set_sr_mw   SR_MW(1:NSPEC) = SQRT(MW) of gckpp_Global (filled by the host from the species database); n = NSPEC, or 0 = off
 *   set_het     het[GCKPP_NHET][ncell], cell-fastest, the HetState fields below (commonIncludeVars.H:112-210; logicals as
 *               0.0 / 1.0) for the cells of the NEXT calls: a host array for the host entry points, a device array for the
 *               _device ones; NULL = off.  conc: the concentrations the laws read, needed only by the stand-alone
 *               gckpp_gpu_update_rconst[_device] (the integrate entry points use their own conc_in). */
enum {
  GCKPP_HET_SUNCOS = 0, GCKPP_HET_STRATBOX, GCKPP_HET_SSA_IS_ALK, GCKPP_HET_SSA_IS_ACID, GCKPP_HET_SSC_IS_ALK,
  GCKPP_HET_SSC_IS_ACID, GCKPP_HET_F_ALK_SSA, GCKPP_HET_F_ALK_SSC, GCKPP_HET_F_ACID_SSA, GCKPP_HET_F_ACID_SSC,
  GCKPP_HET_CLEARFR, GCKPP_HET_ACLAREA, GCKPP_HET_ACLRADI, GCKPP_HET_CL_CONC_SSA, GCKPP_HET_CL_CONC_SSC,
  GCKPP_HET_GAMMA_HO2, GCKPP_HET_H_PLUS, GCKPP_HET_NO3_MOLAL, GCKPP_HET_SO4_MOLAL, GCKPP_HET_HSO4_MOLAL,
  GCKPP_HET_XAREA = 20,   /* xArea(1:14): DU1..DU7, SUL, BKC, ORC, SSA, SSC, SLA, IIC */
  GCKPP_HET_XRADI = 34,   /* xRadi(1:14) */
  /* second part (read only when gckpp_gpu_set_species_data was called): the cloud / halogen laws */
  GCKPP_HET_NATSURFACE = 48, GCKPP_HET_TURNOFFHETRATES, GCKPP_HET_CLDFR, GCKPP_HET_AICE, GCKPP_HET_ALIQ, GCKPP_HET_RICE,
  GCKPP_HET_RLIQ, GCKPP_HET_PHCLOUD, GCKPP_HET_PHSSA /* (1:2) */, GCKPP_HET_CL_CONC_CLD = 58, GCKPP_HET_BR_CONC_CLD,
  GCKPP_HET_BR_CONC_SSA, GCKPP_HET_BR_CONC_SSC, GCKPP_HET_BR_OVER_CL_CLD, GCKPP_HET_BR_OVER_CL_SSA, GCKPP_HET_BR_OVER_CL_SSC,
  GCKPP_HET_FRAC_BR_CLDA, GCKPP_HET_FRAC_BR_CLDC, GCKPP_HET_FRAC_BR_CLDG, GCKPP_HET_FRAC_CL_CLDA, GCKPP_HET_FRAC_CL_CLDC,
  GCKPP_HET_FRAC_CL_CLDG, GCKPP_HET_FRAC_SALACL, GCKPP_HET_FRAC_HSO3_AQ, GCKPP_HET_HSO3M, GCKPP_HET_HCL_THETA,
  GCKPP_HET_HBR_THETA, GCKPP_HET_HNO3_THETA, GCKPP_HET_H_CONC_LCL, GCKPP_HET_H_CONC_SSA, GCKPP_HET_H_CONC_SSC,
  GCKPP_HET_HSO3_AQ, GCKPP_HET_SO3_AQ, GCKPP_HET_TSO3_AQ, GCKPP_HET_AWATER /* (1:2) */, GCKPP_HET_KHETI_SLA = 85 /* (1:11) */,
  /* what N2O5_InorgOrg reads: AClVol, xVol(ORC), xVol(SSC), xH2O(SUL), xH2O(ORC), xH2O(SSC), OMOC_POA, OMOC_OPOA */
  GCKPP_HET_ACLVOL = 96, GCKPP_HET_XVOL_ORC, GCKPP_HET_XVOL_SSC, GCKPP_HET_XH2O_SUL, GCKPP_HET_XH2O_ORC, GCKPP_HET_XH2O_SSC,
  GCKPP_HET_OMOC_POA, GCKPP_HET_OMOC_OPOA,
  GCKPP_NHET = 104
};
int gckpp_gpu_set_sr_mw(gckpp_gpu_handle_t *handle, int n, const double *sr_mw);
/* set_sr_mw plus MW(1:NSPEC) [g/mol] and the Henry's-law constants HENRY_K0 [M/atm], HENRY_CR [K] of gckpp_Global: enables
 * the second part -- 38 more constants (BrNO3, ClNO2, ClNO3, HOBr, HOCl, IONO2, N2O5, O3 + bromide, NO2 / NO3 uptake and cloud
 * loss, NO3 on sea-salt chloride, N2O5 on aerosol / in cloud / + stratospheric HCl; fullchem_RateLawFuncs.F90:803-3238).  14 constants
 * then remain external: K_MT(6), K_CLD(6) and the two HSO3m / SO3mm sums
 * that add the sulfur module's SRHOCl / SRHOBr. */
int gckpp_gpu_set_species_data(gckpp_gpu_handle_t *handle, int n, const double *sr_mw, const double *mw,
                               const double *henry_k0, const double *henry_cr);
int gckpp_gpu_set_het(gckpp_gpu_handle_t *handle, const double *het, const double *conc);

/* Fun(V,F,RCT,Vdot,Aout) over cells (RxnRate diagnostics, fullchem_mod.F90:967-992):
 * vdot [NVAR][ncell] and aout [NREACT][ncell]; either output may be NULL.  Host pointers. */
int gckpp_gpu_fun(gckpp_gpu_handle_t *handle, int ncell, const double *conc, const double *rconst,
                  double *vdot, double *aout);

/* The pieces of Do_FullChem around the integration, so that the chemical state can stay on the device for a whole
 * chemistry step (`_device`: conc / rconst / outputs are device pointers; the others stage host arrays).  The small
 * index lists are always host pointers.  conc is [NSPEC][ncell], cell-fastest, modified in place.
 *   zero_species    GeosCore/fullchem_mod.F90:941-946: C(PL_Kpp_Id(F)) = 0 for the n prod/loss family species ids0 (0-based)
 *   post_integrate  :1284-1287 fullchem_ConvertEquivToAlk (KPP/fullchem/fullchem_SulfurChemFuncs.F90:94-105):
 *                   C(scale_ids0[k]) = C / scale_div[k] (the host passes MW * 7.0e-5; nscale = 0 when sulfate_mod does it);
 *                   :1326-1348 for every species with spc_mask[s] != 0 (Map_KppSpc > 0; NULL = all): negatives[cell] +=
 *                   1 per negative concentration (State_Diag%KppNegatives, REAL*4, may be NULL), then C = MAX(C, 0)
 *   prod_loss       :1463-1492: out[s][cell] = C(ids0[s]) / dt for the nslots prod or loss slots
 *   oh_reactivity   KPP/fullchem/gckpp_Util.F90:983-1040 Get_OHreactivity(C, RCONST) -> ohreact[ncell] (fullchem, Hg) */
int gckpp_gpu_zero_species(gckpp_gpu_handle_t *handle, int ncell, double *conc, int n, const int32_t *ids0);
int gckpp_gpu_zero_species_device(gckpp_gpu_handle_t *handle, int ncell, double *conc, int n, const int32_t *ids0);
int gckpp_gpu_post_integrate(gckpp_gpu_handle_t *handle, int ncell, double *conc, int nscale, const int32_t *scale_ids0,
                             const double *scale_div, const uint8_t *spc_mask, float *negatives);
int gckpp_gpu_post_integrate_device(gckpp_gpu_handle_t *handle, int ncell, double *conc, int nscale, const int32_t *scale_ids0,
                                    const double *scale_div, const uint8_t *spc_mask, float *negatives);
int gckpp_gpu_prod_loss(gckpp_gpu_handle_t *handle, int ncell, const double *conc, double dt, int nslots,
                        const int32_t *ids0, double *out);
int gckpp_gpu_prod_loss_device(gckpp_gpu_handle_t *handle, int ncell, const double *conc, double dt, int nslots,
                               const int32_t *ids0, double *out);
int gckpp_gpu_oh_reactivity(gckpp_gpu_handle_t *handle, int ncell, const double *conc, const double *rconst, double *ohreact);
int gckpp_gpu_oh_reactivity_device(gckpp_gpu_handle_t *handle, int ncell, const double *conc, const double *rconst,
                                   double *ohreact);

/* Pieces exposed for parity tests (host pointers, cell-fastest):
 *   jac:    Jac_SP  -> jvs [LU_NONZERO][ncell]
 *   decomp: KppDecomp in place on jvs; ier[ncell] = 0 or the 1-based singular row
 *   solve:  KppSolve in place on x [NVAR][ncell] */
int gckpp_gpu_jac(gckpp_gpu_handle_t *handle, int ncell, const double *conc, const double *rconst, double *jvs);
int gckpp_gpu_decomp(gckpp_gpu_handle_t *handle, int ncell, double *jvs, int32_t *ier);
int gckpp_gpu_solve(gckpp_gpu_handle_t *handle, int ncell, const double *jvs, double *x);

/* Run this handle's work on a caller-owned CUDA stream (cudaStream_t passed as void*; NULL
 * restores the handle's own stream).  Lets the host order the call against its own transfers
 * and time it with its own events. */
int gckpp_gpu_set_stream(gckpp_gpu_handle_t *handle, void *cuda_stream);

/* Measure the device's FP64 FMA peak [TFLOP/s] with a register-resident DFMA chain kernel on
 * every SM (the non-tensor FP64 pipe is the compute roofline of this path; MEASURED_PEAKS.json
 * only holds HBM and BF16).  ms_out (may be NULL) receives the kernel time. */
int gckpp_gpu_fp64_peak(int device, double *tflops_out, double *ms_out);

/* Statistics of the last integrate call on this handle:
 * stats[0] kernel time of the integrator [ms] (CUDA events on the handle's stream),
 * stats[1] kernel time of Update_RCONST [ms], stats[2] H2D+D2H time [ms] (host entry only),
 * stats[3] cells integrated, stats[4] cells retried, stats[5] cells failed twice,
 * stats[6] kernels launched, stats[7] sum of Nstp, stats[8] sum of Nacc,
 * stats[9] device time of the whole call (rate constants + integration + retry) [ms]. */
int gckpp_gpu_last_stats(gckpp_gpu_handle_t *handle, double *stats /* [16] */);

/* keepActive / keepSpcActive of the auto-reduce solver (gckpp_Global; set by fullchem_AutoReduce_SetKeepActive
 * and fullchem_AutoReduce_KeepHalogensActive, fullchem_AutoReduceFuncs.F90:40-140): n 0-based variable-species
 * indices that are never removed from the implicit system; n = 0 switches keepActive off.
 * Auto-reduce itself is selected like in the reference: ICNTRL(12)=1, threshold RCNTRL(12) (default 100) or,
 * with a target species ICNTRL(14) > 0, RCNTRL(14) * max(LossY, Prod) of that species; it runs on the
 * table-driven kernel (ros_yIntegrator, gckpp_Integrator.F90:789-1237); rstatus[3] returns the threshold.
 * The append variant (ICNTRL(13)=1) is not available (-12). */
int gckpp_gpu_set_keep_active(gckpp_gpu_handle_t *handle, int n, const int32_t *idx0);

/* Host-only description of the shared-memory kernel's static plan for a mechanism (no GPU needed):
 * info[0] dynamic shared memory per block [bytes], info[1] streamed table rows (512 B) per attempt,
 * info[2] resident table rows, info[3] rounds in the directory, info[4..6] LU / forward / backward
 * rounds, info[7] cells per block. */
int gckpp_gpu_plan_info(int mech_id, int32_t *info /* [8] */);
/* Test hook: the per-warp table streams of the warp-group integrator (csrc/ros_warp.h) exactly as the host plan
 * lays them out, so that CPU tests can replay them (tests/test_wsched.py).  No reference counterpart. */
int gckpp_gpu_warp_plan(int mech_id, int32_t *info, int info_cap, uint32_t *stream_out, int64_t stream_cap_words);

/* Last error text (thread-local). */
const char *gckpp_gpu_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
