// C ABI of the B200 KPP chemistry path (see include/gckpp_gpu.h for the contract and the
// reference interfaces each entry point replaces).  Host side only: option decoding
// (= Rosenbrock(), gckpp_Integrator.F90:165-531), buffer management, kernel launches.
// There is NO CPU fallback: every entry point fails loudly if CUDA is unavailable.
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/gckpp_gpu.h"
#include "ros_common.cuh"
#include "kernels.h"
#include "ros_smem.h"
#include "ros_warp.h"
#include "ros_lane.h"

#include "gen/fullchem_tables.h"
#include "gen/Hg_tables.h"
#include "gen/carbon_tables.h"
#include "gen/fullchem_sched.h"
#include "gen/Hg_sched.h"
#include "gen/fullchem_wsched.h"
#include "gen/Hg_wsched.h"
#include "gen/fullchem_lsched.h"
#include "gen/Hg_lsched.h"
#include "gen/fullchem_names.h"
#include "gen/Hg_names.h"
#include "gen/carbon_names.h"

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
#define CUDA_TRY(x)                                                                      \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess)                                                               \
      return fail(-(1000 + (int)e_), "%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
  } while (0)

static const gckpp_host_tables_t *host_tables(int mech_id)
{
  switch (mech_id) {
  case GCKPP_MECH_FULLCHEM: return &fullchem_tables;
  case GCKPP_MECH_HG: return &Hg_tables;
  case GCKPP_MECH_CARBON: return &carbon_tables;
  default: return nullptr;
  }
}

static const gckpp_sched_tables_t *host_sched(int mech_id)
{
  switch (mech_id) {
  case GCKPP_MECH_FULLCHEM: return &fullchem_sched;
  case GCKPP_MECH_HG: return &Hg_sched;
  default: return nullptr;
  }
}

static const gckpp_wsched_tables_t *host_wsched(int mech_id)
{
  switch (mech_id) {
  case GCKPP_MECH_FULLCHEM: return &fullchem_wsched;
  case GCKPP_MECH_HG: return &Hg_wsched;
  default: return nullptr;
  }
}

static const gckpp_lsched_tables_t *host_lsched(int mech_id)
{
  switch (mech_id) {
  case GCKPP_MECH_FULLCHEM: return &fullchem_lsched;
  case GCKPP_MECH_HG: return &Hg_lsched;
  default: return nullptr;
  }
}

struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  int ensure(size_t n)
  {
    if (n <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) return (int)e;
    bytes = n;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <class T> T *as() { return (T *)p; }
};

struct WaveSlot;
static void free_slots(struct gckpp_gpu_handle *h);
struct gckpp_gpu_handle {
  int mech_id = 0, device = 0, max_cells = 0;
  const gckpp_host_tables_t *T = nullptr;
  MechDev M{};
  std::vector<void *> table_allocs;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  cudaEvent_t ev[8]{};
  int sm_count = 0;
  // integrator launch geometry + workspace
  int threads = 128, blocks_per_sm = 5, max_blocks = 0;
  WsLayout L{};
  DevBuf work, next, sums, tol, cell_list, counter, rconst_work, scratch;
  // staging for the host entry points
  DevBuf s_conc_in, s_conc_out, s_rconst, s_met, s_photol, s_khet, s_hstart, s_active, s_ist, s_rst, s_ierr;
  int opt_retry = 0, opt_kernel = -1, opt_wave_cells = 0, opt_pin = 0, opt_dev_wave = 0;
  // shared-memory kernel: host plan + device copies of its tables
  int sm_ready = 0, sm_blocks_cap = 0;
  SmemHostPlan plan;
  SmemArgs sargs{};
  DevBuf sm_rcs, sm_scr, sm_stream, sm_res, sm_boff, sm_dir, sm_tpos, sm_crow, sm_aw, sm_bw, sm_coefs, sm_diag, sm_rowcol, ar_mask;
  int last_kernel = 0;
  DevBuf sm_uscale;
  // warp-per-cell kernel: host plan + device copies of its tables
  int w_ready = 0;
  WarpHostPlan wplan;
  WarpArgs wargs{};
  DevBuf w_stream, w_aw, w_bw, w_diag, w_tpos, w_coefs, w_rcs;
  // lane kernel: device copies of its stream tables, workspace of the resident blocks
  int l_ready = 0, l_blocks = 0;
  LaneArgs largs{};
  DevBuf l_tab[7], l_lit, l_ws;
  const double *ohr_coef = nullptr; const int *ohr_rxn = nullptr, *ohr_spc = nullptr;   // Get_OHreactivity terms (device)
  DevBuf small;                            // the little index lists of the post-integrate entry points
  // heterogeneous laws on the device: SR_MW (device copy), the caller's HetState fields and, for the stand-alone
  // Update_RCONST entry points, its concentrations (host or device pointers, whatever the next call takes)
  DevBuf srmw; int srmw_n = 0, spc_data_n = 0;      // [SR_MW | MW | HENRY_K0 | HENRY_CR]
  const double *het_user = nullptr, *het_conc_user = nullptr;
  DevBuf keep_spc; int keep_n = 0;         // keepSpcActive of the auto-reduce solver
  // pipelined host entry: copy streams and the identity cell list
  cudaStream_t s_in = nullptr, s_out = nullptr;
  DevBuf ident; int ident_n = 0;
  int opt_chunks = 0;
  struct WaveSlot *slots = nullptr;
  double stats[16]{};
};

extern "C" const char *gckpp_gpu_last_error(void) { return g_err.c_str(); }

extern "C" int gckpp_gpu_dims(int mech_id, int32_t *dims)
{
  const gckpp_host_tables_t *T = host_tables(mech_id);
  if (!T || !dims) return fail(-10, "gckpp_gpu_dims: bad mechanism id %d", mech_id);
  dims[0] = T->nvar; dims[1] = T->nfix; dims[2] = T->nspec; dims[3] = T->nreact;
  dims[4] = T->nnz; dims[5] = T->nphot; dims[6] = T->next;
  return 0;
}

extern "C" const char *gckpp_gpu_spc_name(int mech_id, int i)
{
  const gckpp_host_tables_t *T = host_tables(mech_id);
  if (!T || i < 0 || i >= T->nspec) return nullptr;
  switch (mech_id) {
  case GCKPP_MECH_FULLCHEM: return fullchem_spc_names[i];
  case GCKPP_MECH_HG: return Hg_spc_names[i];
  default: return carbon_spc_names[i];
  }
}

template <class T> static int upload(gckpp_gpu_handle *h, const T *src, size_t n, const T **dst)
{
  void *p = nullptr;
  if (n == 0) n = 1;
  CUDA_TRY(cudaMalloc(&p, n * sizeof(T)));
  h->table_allocs.push_back(p);
  if (src) CUDA_TRY(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
  *dst = (const T *)p;
  return 0;
}

static WsLayout make_layout(const gckpp_host_tables_t *T)
{
  WsLayout L;
  int o = 0;
  L.Y = o;  o += T->nspec;
  L.YN = o; o += T->nspec;
  L.F0 = o; o += T->nvar;
  L.FC = o; o += T->nvar;
  L.K = o;  o += 6 * T->nvar;           // up to 6 stages (Rodas4)
  L.G = o;  o += T->nnz > 0 ? T->nnz : 1;
  L.RC = o; o += T->nreact;
  L.AB = o; o += (T->nreact > T->nb ? T->nreact : T->nb);
  L.W = o;  o += T->nvar;
  L.PR = o; o += T->nvar;              // auto-reduce: initial Prod, Loss and the keep mask
  L.LS = o; o += T->nvar;
  L.MK = o; o += T->nvar;
  L.total = o;
  return L;
}

extern "C" int gckpp_gpu_init(int mech_id, int device, int max_cells, gckpp_gpu_handle_t **out)
{
  if (!out) return fail(-10, "gckpp_gpu_init: handle pointer is NULL");
  *out = nullptr;
  const gckpp_host_tables_t *T = host_tables(mech_id);
  if (!T) return fail(-10, "gckpp_gpu_init: bad mechanism id %d", mech_id);
  if (max_cells <= 0) return fail(-10, "gckpp_gpu_init: max_cells must be positive");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(-10, "gckpp_gpu_init: device %d of %d", device, ndev);
  CUDA_TRY(cudaSetDevice(device));
  gckpp_gpu_handle *h = new gckpp_gpu_handle();
  h->mech_id = mech_id; h->device = device; h->max_cells = max_cells; h->T = T;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  CUDA_TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  for (auto &e : h->ev) CUDA_TRY(cudaEventCreate(&e));
  MechDev &M = h->M;
  M.nvar = T->nvar; M.nfix = T->nfix; M.nspec = T->nspec; M.nreact = T->nreact; M.nnz = T->nnz;
  M.nb = T->nb; M.nphot = T->nphot; M.next = T->next; M.fun_split = T->fun_split;
  int rc;
#define UP(field, src, n, type) if ((rc = upload<type>(h, (const type *)(src), (size_t)(n), (const type **)&M.field))) { gckpp_gpu_finalize(h); return rc; }
  UP(a_term, T->a_term, T->nreact, int4);
  UP(p_ptr, T->p_ptr, T->nvar + 1, int); UP(p_coef, T->p_coef, T->np, double); UP(p_rxn, T->p_rxn, T->np, int);
  UP(d_ptr, T->d_ptr, T->nvar + 1, int); UP(d_term, T->d_term, T->nd, int4);
  UP(v_ptr, T->v_ptr, T->nvar + 1, int); UP(v_coef, T->v_coef, T->nv, double); UP(v_rxn, T->v_rxn, T->nv, int);
  if (T->nnz > 0) {
    UP(crow, T->crow, T->nvar + 1, int); UP(diag, T->diag, T->nvar, int); UP(icol, T->icol, T->nnz, int);
    UP(b_term, T->b_term, T->nb, int4);
    UP(j_ptr, T->j_ptr, T->nnz + 1, int); UP(j_coef, T->j_coef, T->nj, double); UP(j_b, T->j_b, T->nj, int);
  }
  UP(lit, T->lit, T->nlit, double);
  if (T->nohr > 0) {
    if ((rc = upload<double>(h, T->ohr_coef, (size_t)T->nohr, &h->ohr_coef)) || (rc = upload<int>(h, T->ohr_rxn, (size_t)T->nohr, &h->ohr_rxn)) ||
        (rc = upload<int>(h, T->ohr_spc, (size_t)T->nohr, &h->ohr_spc))) { gckpp_gpu_finalize(h); return rc; }
  }
#undef UP
  h->L = make_layout(T);
  h->max_blocks = h->sm_count * h->blocks_per_sm;
  if (h->next.ensure(sizeof(int)) || h->sums.ensure(128 * sizeof(unsigned long long)) ||
      h->tol.ensure(2 * sizeof(double) * T->nvar) || h->counter.ensure(4 * sizeof(int))) {
    gckpp_gpu_finalize(h);
    return fail(-1002, "gckpp_gpu_init: out of device memory");
  }
  *out = h;
  return 0;
}

extern "C" int gckpp_gpu_finalize(gckpp_gpu_handle_t *h)
{
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (void *p : h->table_allocs) cudaFree(p);
  DevBuf *bufs[] = {&h->work, &h->next, &h->sums, &h->tol, &h->cell_list, &h->counter, &h->rconst_work, &h->scratch,
                    &h->s_conc_in, &h->s_conc_out, &h->s_rconst, &h->s_met, &h->s_photol, &h->s_khet, &h->s_hstart,
                    &h->s_active, &h->s_ist, &h->s_rst, &h->s_ierr,
                    &h->w_stream, &h->w_aw, &h->w_bw, &h->w_diag, &h->w_tpos, &h->w_coefs, &h->w_rcs,
                    &h->l_tab[0], &h->l_tab[1], &h->l_tab[2], &h->l_tab[3], &h->l_tab[4], &h->l_tab[5], &h->l_tab[6], &h->l_lit, &h->l_ws,
                    &h->small, &h->srmw,
                    &h->keep_spc, &h->sm_uscale, &h->ident, &h->sm_rcs, &h->sm_scr, &h->sm_stream, &h->sm_res, &h->sm_boff, &h->sm_dir, &h->sm_tpos, &h->sm_crow, &h->sm_aw, &h->sm_bw, &h->sm_coefs, &h->sm_diag, &h->sm_rowcol, &h->ar_mask};
  for (DevBuf *b : bufs) b->release();
  free_slots(h);
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  delete h;
  return 0;
}

extern "C" int gckpp_gpu_set_stream(gckpp_gpu_handle_t *h, void *cuda_stream)
{
  if (!h) return fail(-10, "NULL handle");
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return 0;
}

extern "C" int gckpp_gpu_fp64_peak(int device, double *tflops_out, double *ms_out)
{
  if (!tflops_out) return fail(-10, "NULL argument");
  CUDA_TRY(cudaSetDevice(device));
  double tf = 0, ms = 0;
  CUDA_TRY(measure_fp64_peak(&tf, &ms));
  *tflops_out = tf;
  if (ms_out) *ms_out = ms;
  return 0;
}

extern "C" int gckpp_gpu_set_option(gckpp_gpu_handle_t *h, const char *key, int value)
{
  if (!h || !key) return fail(-10, "gckpp_gpu_set_option: NULL argument");
  if (!strcmp(key, "retry")) h->opt_retry = value;
  else if (!strcmp(key, "kernel")) h->opt_kernel = value;
  else if (!strcmp(key, "wave_cells")) { if (value < 0) return fail(-10, "wave_cells out of range"); h->opt_wave_cells = value; }
  else if (!strcmp(key, "pin")) h->opt_pin = value;
  else if (!strcmp(key, "device_wave_cells")) { if (value < 0) return fail(-10, "device_wave_cells out of range"); h->opt_dev_wave = value; }
  else if (!strcmp(key, "blocks_per_sm")) { if (value < 1 || value > 16) return fail(-10, "blocks_per_sm out of range"); h->blocks_per_sm = value; h->max_blocks = h->sm_count * value; }
  else if (!strcmp(key, "blocks_cap")) { h->sm_blocks_cap = value; }
  else if (!strcmp(key, "chunks")) { if (value < 0 || value > 4096) return fail(-10, "chunks out of range"); h->opt_chunks = value; }
  else if (!strcmp(key, "threads")) { if (value < 32 || value > 1024 || value % 32) return fail(-10, "threads must be a multiple of 32"); h->threads = value; }
  else return fail(-10, "gckpp_gpu_set_option: unknown option '%s'", key);
  return 0;
}

// ---- Rosenbrock(): option decoding (gckpp_Integrator.F90:165-531 after Integrate's merge :100-117)
static void set_method(RosOpts &o, int id)
{
  memset(o.A, 0, sizeof o.A); memset(o.C, 0, sizeof o.C); memset(o.M, 0, sizeof o.M);
  memset(o.E, 0, sizeof o.E); memset(o.Alpha, 0, sizeof o.Alpha); memset(o.Gamma, 0, sizeof o.Gamma);
  memset(o.NewF, 0, sizeof o.NewF);
  switch (id) {
  case 1: {  // Ros2 (:2062-2105)
    double g = 1.0 + 1.0 / sqrt(2.0);
    o.S = 2; o.A[0] = 1.0 / g; o.C[0] = -2.0 / g; o.NewF[0] = o.NewF[1] = 1;
    o.M[0] = 3.0 / (2.0 * g); o.M[1] = 1.0 / (2.0 * g); o.E[0] = 1.0 / (2.0 * g); o.E[1] = 1.0 / (2.0 * g);
    o.ELO = 2.0; o.Alpha[1] = 1.0; o.Gamma[0] = g; o.Gamma[1] = -g;
    break; }
  case 2:    // Ros3 (:2108-2157)
    o.S = 3; o.A[0] = 1.0; o.A[1] = 1.0;
    o.C[0] = -0.10156171083877702091975600115545e+01; o.C[1] = 0.40759956452537699824805835358067e+01;
    o.C[2] = 0.92076794298330791242156818474003e+01;
    o.NewF[0] = o.NewF[1] = 1;
    o.M[0] = 0.1e+01; o.M[1] = 0.61697947043828245592553615689730e+01; o.M[2] = -0.42772256543218573326238373806514;
    o.E[0] = 0.5; o.E[1] = -0.29079558716805469821718236208017e+01; o.E[2] = 0.22354069897811569627360909276199;
    o.ELO = 3.0;
    o.Alpha[1] = o.Alpha[2] = 0.43586652150845899941601945119356;
    o.Gamma[0] = 0.43586652150845899941601945119356; o.Gamma[1] = 0.24291996454816804366592249683314;
    o.Gamma[2] = 0.21851380027664058511513169485832e+01;
    break;
  case 3:    // Ros4 (:2164-2231)
    o.S = 4;
    o.A[0] = 0.2000000000000000e+01; o.A[1] = 0.1867943637803922e+01; o.A[2] = 0.2344449711399156;
    o.A[3] = o.A[1]; o.A[4] = o.A[2];
    o.C[0] = -0.7137615036412310e+01; o.C[1] = 0.2580708087951457e+01; o.C[2] = 0.6515950076447975;
    o.C[3] = -0.2137148994382534e+01; o.C[4] = -0.3214669691237626; o.C[5] = -0.6949742501781779;
    o.NewF[0] = o.NewF[1] = o.NewF[2] = 1;
    o.M[0] = 0.2255570073418735e+01; o.M[1] = 0.2870493262186792; o.M[2] = 0.4353179431840180; o.M[3] = 0.1093502252409163e+01;
    o.E[0] = -0.2815431932141155; o.E[1] = -0.7276199124938920e-01; o.E[2] = -0.1082196201495311; o.E[3] = -0.1093502252409163e+01;
    o.ELO = 4.0;
    o.Alpha[1] = 0.1145640000000000e+01; o.Alpha[2] = o.Alpha[3] = 0.6552168638155900;
    o.Gamma[0] = 0.5728200000000000; o.Gamma[1] = -0.1769193891319233e+01; o.Gamma[2] = 0.7592633437920482;
    o.Gamma[3] = -0.1049021087100450;
    break;
  case 5:    // Rodas4 (:2310-2405)
    o.S = 6;
    o.Alpha[1] = 0.386; o.Alpha[2] = 0.210; o.Alpha[3] = 0.630; o.Alpha[4] = o.Alpha[5] = 1.0;
    o.Gamma[0] = 0.25; o.Gamma[1] = -0.1043; o.Gamma[2] = 0.1035; o.Gamma[3] = -0.3620000000000023e-01;
    o.A[0] = 0.1544000000000000e+01; o.A[1] = 0.9466785280815826; o.A[2] = 0.2557011698983284;
    o.A[3] = 0.3314825187068521e+01; o.A[4] = 0.2896124015972201e+01; o.A[5] = 0.9986419139977817;
    o.A[6] = 0.1221224509226641e+01; o.A[7] = 0.6019134481288629e+01; o.A[8] = 0.1253708332932087e+02;
    o.A[9] = -0.6878860361058950; o.A[10] = o.A[6]; o.A[11] = o.A[7]; o.A[12] = o.A[8]; o.A[13] = o.A[9]; o.A[14] = 1.0;
    o.C[0] = -0.5668800000000000e+01; o.C[1] = -0.2430093356833875e+01; o.C[2] = -0.2063599157091915;
    o.C[3] = -0.1073529058151375; o.C[4] = -0.9594562251023355e+01; o.C[5] = -0.2047028614809616e+02;
    o.C[6] = 0.7496443313967647e+01; o.C[7] = -0.1024680431464352e+02; o.C[8] = -0.3399990352819905e+02;
    o.C[9] = 0.1170890893206160e+02; o.C[10] = 0.8083246795921522e+01; o.C[11] = -0.7981132988064893e+01;
    o.C[12] = -0.3152159432874371e+02; o.C[13] = 0.1631930543123136e+02; o.C[14] = -0.6058818238834054e+01;
    o.M[0] = o.A[6]; o.M[1] = o.A[7]; o.M[2] = o.A[8]; o.M[3] = o.A[9]; o.M[4] = 1.0; o.M[5] = 1.0;
    o.E[5] = 1.0;
    for (int i = 0; i < 6; i++) o.NewF[i] = 1;
    o.ELO = 4.0;
    break;
  case 6:    // Rang3 (:2412-2476)
    o.S = 4;
    o.A[0] = 5.09052051067020e+00; o.A[1] = 5.09052051067020e+00; o.A[2] = 0.0;
    o.A[3] = 4.97628111010787e+00; o.A[4] = 2.77268164715849e-02; o.A[5] = 2.29428036027904e-01;
    o.C[0] = -1.16790812312283e+01; o.C[1] = -1.64057326467367e+01; o.C[2] = -2.77268164715850e-01;
    o.C[3] = -8.38103960500476e+00; o.C[4] = -8.48328409199343e-01; o.C[5] = 2.87009860433106e-01;
    o.M[0] = 5.22582761233094e+00; o.M[1] = -5.56971148154165e-01; o.M[2] = 3.57979469353645e-01; o.M[3] = 1.72337398521064e+00;
    o.E[0] = -5.16845212784040e+00; o.E[1] = -1.26351942603842e+00; o.E[2] = -1.11022302462516e-16; o.E[3] = 2.22044604925031e-16;
    o.Alpha[1] = 2.21878746765329e+00; o.Alpha[2] = 2.21878746765329e+00; o.Alpha[3] = 1.55392337535788e+00;
    o.Gamma[0] = 4.35866521508459e-01; o.Gamma[1] = -1.78292094614483e+00; o.Gamma[2] = -2.46541900496934e+00;
    o.Gamma[3] = -8.05529997906370e-01;
    for (int i = 0; i < 4; i++) o.NewF[i] = 1;
    o.ELO = 3.0;
    break;
  default:   // 0, 4: Rodas3 (:2239-2303) -- the method GEOS-Chem selects
    o.S = 4;
    o.A[1] = 2.0; o.A[3] = 2.0; o.A[5] = 1.0;
    o.C[0] = 4.0; o.C[1] = 1.0; o.C[2] = -1.0; o.C[3] = 1.0; o.C[4] = -1.0; o.C[5] = -(8.0 / 3.0);
    o.NewF[0] = 1; o.NewF[2] = 1; o.NewF[3] = 1;
    o.M[0] = 2.0; o.M[2] = 1.0; o.M[3] = 1.0;
    o.E[3] = 1.0;
    o.ELO = 3.0;
    o.Alpha[2] = 1.0; o.Alpha[3] = 1.0;
    o.Gamma[0] = 0.5; o.Gamma[1] = 1.5;
    break;
  }
}

struct Decoded {
  RosOpts o;
  int ICNTRL[20];
  double RCNTRL[20];
  int autoreduce;
};

// returns 0 or the reference's IERR (-1..-5); -12 = unsupported option combination
static int decode_options(const gckpp_host_tables_t *T, double tin, double tout, const int32_t *icntrl_u,
                          const double *rcntrl_u, const double *atol, const double *rtol, Decoded &d)
{
  int *IC = d.ICNTRL;
  double *RC = d.RCNTRL;
  for (int i = 0; i < 20; i++) { IC[i] = 0; RC[i] = 0.0; }
  IC[14] = 5;
  if (icntrl_u) for (int i = 0; i < 20; i++) if (icntrl_u[i] != 0) IC[i] = icntrl_u[i];
  if (rcntrl_u) for (int i = 0; i < 20; i++) if (rcntrl_u[i] > 0) RC[i] = rcntrl_u[i];
  if (IC[14] != -1) return fail(-12, "ICNTRL(15) must be -1: rates are not refreshed inside the integrator (got %d)", IC[14]);
  RosOpts &o = d.o;
  o.Autonomous = !(IC[0] == 0);
  o.VectorTol = (IC[1] == 0);
  int UplimTol = o.VectorTol ? T->nvar : 1;
  if (IC[2] < 0 || IC[2] > 6) return fail(-2, "Selected Rosenbrock method not implemented: ICNTRL(3)=%d", IC[2]);
  set_method(o, IC[2]);
  if (IC[3] == 0) o.Max_no_steps = 200000;
  else if (IC[3] > 0) o.Max_no_steps = IC[3];
  else return fail(-1, "Improper value for maximal no of steps: ICNTRL(4)=%d", IC[3]);
  d.autoreduce = (IC[11] == 1);
  o.ClipNegative = (IC[15] == 1);
  {  // WLAMCH('E') (gckpp_LinearAlgebra.F90:3861-3895) = 2^-52
    o.Roundoff = DBL_EPSILON;
  }
  o.Hmin = RC[0];                       // RCNTRL > 0 merged above, negative values cannot reach here
  o.Hmax = (RC[1] == 0.0) ? fabs(tout - tin) : fmin(fabs(RC[1]), fabs(tout - tin));
  o.Hstart_rcntrl = RC[2];
  o.FacMin = (RC[3] == 0.0) ? 0.2 : RC[3];
  o.FacMax = (RC[4] == 0.0) ? 6.0 : RC[4];
  o.FacRej = (RC[5] == 0.0) ? 0.1 : RC[5];
  o.FacSafe = (RC[6] == 0.0) ? 0.9 : RC[6];
  if (T->nnz > 0) {
    if (!atol || !rtol) return fail(-10, "atol/rtol are required");
    for (int i = 0; i < UplimTol; i++)
      if ((atol[i] <= 0.0) || (rtol[i] <= 10.0 * o.Roundoff) || (rtol[i] >= 1.0))
        return fail(-5, "Improper tolerance values: AbsTol(%d)=%g RelTol(%d)=%g", i + 1, atol[i], i + 1, rtol[i]);
  }
  o.Tstart = tin; o.Tend = tout;
  o.Direction = (tout >= tin) ? +1 : -1;
  return 0;
}

// ---- device-side helpers implemented in kernels_misc.cu
static int ensure_workspace(gckpp_gpu_handle *h, int blocks)
{
  size_t nwarps = (size_t)blocks * (h->threads / 32);
  size_t stride = (size_t)h->L.total * 32;
  size_t bytes = nwarps * stride * sizeof(double);
  if (bytes > h->work.bytes) {
    if (h->work.ensure(bytes)) return fail(-1002, "out of device memory for the integrator workspace (%zu MB)", bytes >> 20);
    cudaMemsetAsync(h->work.p, 0, bytes, h->stream);
  }
  return 0;
}

// Upload the tables of the shared-memory kernel once per handle.
static int prepare_smem(gckpp_gpu_handle *h)
{
  if (h->sm_ready) return 0;
  const gckpp_sched_tables_t *S = host_sched(h->mech_id);
  if (!S || !smem_kernel_supports(h->mech_id)) return fail(-11, "shared-memory kernel not available for this mechanism");
  int prc = smem_plan_build(h->mech_id, h->T, S, h->plan);
  if (prc) return fail(-11, "shared-memory kernel: plan failed (%d)", prc);
  SmemHostPlan &p = h->plan;
  int maxsm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  if (maxsm < p.s_total + 64) return fail(-11, "shared-memory kernel needs %d bytes of shared memory, device offers %d", p.s_total, maxsm);
  struct Up { DevBuf *b; const void *src; size_t bytes; };
  Up ups[] = {
    {&h->sm_stream, p.stream.data(), p.stream.size() * 4}, {&h->sm_res, p.resident.data(), p.resident.size() * 4},
    {&h->sm_boff, p.boff.data(), p.boff.size() * 2},       {&h->sm_dir, p.dir.data(), p.dir.size() * 4},
    {&h->sm_tpos, S->tpos, 32 * 32 * 2},                   {&h->sm_diag, p.diag.data(), p.diag.size() * 2},
    {&h->sm_crow, p.crow.data(), p.crow.size() * 2},       {&h->sm_aw, p.aw.data(), p.aw.size() * 4},
    {&h->sm_bw, p.bw.data(), p.bw.size() * 4},             {&h->sm_uscale, p.uscale.data(), p.uscale.size() * 4},             {&h->sm_coefs, S->coefs, sizeof(double) * (size_t)S->ncoef},
    {&h->sm_rowcol, p.rowcol.data(), p.rowcol.size() * 4},
  };
  for (Up &u : ups) {
    if (u.b->ensure(u.bytes ? u.bytes : 16)) return fail(-1002, "out of device memory for the kernel tables");
    if (u.bytes) CUDA_TRY(cudaMemcpy(u.b->p, u.src, u.bytes, cudaMemcpyHostToDevice));
  }
  SmemArgs &A = h->sargs;
  A.stream = h->sm_stream.as<uint4>();
  for (int w = 0; w < SMEM_NW; w++) { A.warp_off[w] = p.warp_off[w]; A.warp_rows[w] = p.warp_rows[w]; }
  A.resident = h->sm_res.as<uint4>(); A.res_rows = (int)(p.resident.size() / 128);
  A.boff = h->sm_boff.as<uint16_t>(); A.nresb = (int)p.boff.size();
  A.dir = h->sm_dir.as<uint32_t>(); A.ndir = (int)p.dir.size();
  A.o_lu = p.o_lu; A.n_lu = p.n_lu; A.o_fwd = p.o_fwd; A.n_fwd = p.n_fwd; A.o_bwd = p.o_bwd; A.n_bwd = p.n_bwd; A.o_fwd1 = p.o_fwd1;
  A.tpos = h->sm_tpos.as<uint16_t>();
  A.diag = h->sm_diag.as<uint16_t>(); A.crow = h->sm_crow.as<uint16_t>();
  A.aw = h->sm_aw.as<uint32_t>(); A.bw = h->sm_bw.as<uint32_t>();
  A.coefs = h->sm_coefs.as<double>();
  A.uscale = h->sm_uscale.as<uint32_t>(); A.nuscale = (int)p.uscale.size();
  A.lit = h->M.lit;
  if (h->sm_rcs.ensure(sizeof(double) * smem_rcs_doubles_per_block(h->mech_id) * (size_t)h->sm_count)) return fail(-1002, "out of device memory");
  A.rcs = h->sm_rcs.as<double>();
  if (h->sm_scr.ensure(sizeof(double) * smem_scr_doubles_per_block(h->mech_id) * (size_t)h->sm_count)) return fail(-1002, "out of device memory");
  A.scr = h->sm_scr.as<double>();
  A.s_res = p.s_res; A.s_tpos = p.s_tpos; A.s_boff = p.s_boff; A.s_dir = p.s_dir; A.s_diag = p.s_diag; A.s_crow = p.s_crow;
  A.s_total = p.s_total;
  h->sm_ready = 1;
  return 0;
}

// Upload the tables of the warp-per-cell kernel once per handle.
static int prepare_warp(gckpp_gpu_handle *h)
{
  if (h->w_ready) return 0;
  const gckpp_wsched_tables_t *S = host_wsched(h->mech_id);
  if (!S || !warp_kernel_supports(h->mech_id)) return fail(-11, "warp-per-cell kernel not available for this mechanism");
  int prc = warp_plan_build(h->mech_id, h->T, S, h->wplan);
  if (prc) return fail(-11, "warp-per-cell kernel: plan failed (%d)", prc);
  WarpHostPlan &p = h->wplan;
  int maxsm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  if (maxsm < warp_smem_bytes(h->mech_id)) return fail(-11, "warp-per-cell kernel needs %d bytes of shared memory, device offers %d", warp_smem_bytes(h->mech_id), maxsm);
  struct Up { DevBuf *b; const void *src; size_t bytes; };
  Up ups[] = {
    {&h->w_stream, p.stream.data(), p.stream.size() * 4}, {&h->w_aw, p.aw.data(), p.aw.size() * 4},
    {&h->w_bw, p.bw.data(), p.bw.size() * 4},             {&h->w_diag, p.diag.data(), p.diag.size() * 2},
    {&h->w_tpos, S->tpos, 32 * 32 * 2},                   {&h->w_coefs, S->coefs, sizeof(double) * (size_t)S->ncoef},
  };
  for (Up &u : ups) {
    if (u.b->ensure(u.bytes ? u.bytes : 16)) return fail(-1002, "out of device memory for the kernel tables");
    if (u.bytes) CUDA_TRY(cudaMemcpy(u.b->p, u.src, u.bytes, cudaMemcpyHostToDevice));
  }
  WarpArgs &A = h->wargs;
  A.stream = h->w_stream.as<uint4>();
  for (int w = 0; w < WARP_WG; w++) {
    A.w_off[w] = p.w_off[w]; A.w_rows[w] = p.w_rows[w];
    for (int i = 0; i < WARP_NSEG; i++) A.seg_off[w][i] = p.seg_off[w][i];
    for (int i = 0; i < WARP_NPH; i++) A.nb[w][i] = p.nb[w][i];
  }
  for (int i = 0; i < WARP_NPH; i++) A.nlev[i] = p.nlev[i];
  A.tpos = h->w_tpos.as<uint16_t>(); A.diag = h->w_diag.as<uint16_t>();
  A.aw = h->w_aw.as<uint32_t>(); A.bw = h->w_bw.as<uint32_t>();
  A.coefs = h->w_coefs.as<double>(); A.lit = h->M.lit;
  size_t ngroups = (size_t)h->sm_count * warp_cells_per_block(h->mech_id);
  if (h->w_rcs.ensure(sizeof(double) * warp_rcs_doubles_per_group(h->mech_id) * ngroups)) return fail(-1002, "out of device memory");
  A.rcs = h->w_rcs.as<double>();
  A.s_total = warp_smem_bytes(h->mech_id);
  h->w_ready = 1;
  return 0;
}

// Which integrator kernel serves this call: 2 = warp-per-cell (default for Rodas3, ICNTRL(3) = 0 or 4, the
// method GEOS-Chem selects), 1 = the block-synchronous shared-memory kernel of round 1 (kept for comparison),
// 0 = the table-driven reference-order kernel (every method, auto-reduce).
// Upload the stream tables of the lane kernel once per handle and size its workspace: one warp per block, as many
// blocks per SM as its shared memory (the VEC of 32 cells + the rings) allows.
static int prepare_lane(gckpp_gpu_handle *h)
{
  if (h->l_ready) return 0;
  const gckpp_lsched_tables_t *S = host_lsched(h->mech_id);
  if (!S) return fail(-11, "lane kernel not available for this mechanism");
  int maxsm = 0, smsm = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
  CUDA_TRY(cudaDeviceGetAttribute(&smsm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device));
  const size_t smem = lane_smem_bytes(S->d);
  if ((size_t)maxsm < smem) return fail(-11, "lane kernel needs %zu bytes of shared memory, device offers %d", smem, maxsm);
  int bps = (int)((size_t)smsm / (smem + 1024));
  if (bps < 1) bps = 1;
  if (bps > 16) bps = 16;
  LaneArgs &A = h->largs;
  A.d = S->d;
  const uint32_t *src[7] = {S->rates_a, S->sums_v, S->rates_b, S->sums_j, S->lu, S->fwd, S->bwd};
  const int nch[7] = {S->nchunk_ra, S->nchunk_sv, S->nchunk_rb, S->nchunk_sj, S->nchunk_lu, S->nchunk_fwd, S->nchunk_bwd};
  const int recq[7] = {2, 4, 2, 4, 2, 2, 2};
  for (int i = 0; i < 7; i++) {
    const size_t chunk = (size_t)16 * recq[i] * 16, bytes = (size_t)nch[i] * chunk;     // + one chunk of zero records
    if (h->l_tab[i].ensure(bytes + chunk)) return fail(-1002, "out of device memory for the kernel tables");
    CUDA_TRY(cudaMemcpy(h->l_tab[i].p, src[i], bytes, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset((char *)h->l_tab[i].p + bytes, 0, chunk));
  }
  if (h->l_lit.ensure(sizeof(double) * (S->d.nlit > 0 ? S->d.nlit : 1))) return fail(-1002, "out of device memory");
  if (S->d.nlit) CUDA_TRY(cudaMemcpy(h->l_lit.p, S->lit, sizeof(double) * S->d.nlit, cudaMemcpyHostToDevice));
  A.rates_a = h->l_tab[0].as<uint4>(); A.sums_v = h->l_tab[1].as<uint4>(); A.rates_b = h->l_tab[2].as<uint4>();
  A.sums_j = h->l_tab[3].as<uint4>(); A.lu = h->l_tab[4].as<uint4>(); A.fwd = h->l_tab[5].as<uint4>(); A.bwd = h->l_tab[6].as<uint4>();
  A.nchunk_ra = nch[0]; A.nchunk_sv = nch[1]; A.nchunk_rb = nch[2]; A.nchunk_sj = nch[3]; A.nchunk_lu = nch[4];
  A.nchunk_fwd = nch[5]; A.nchunk_bwd = nch[6];
  A.lit = h->l_lit.as<double>();
  const LaneDims &d = S->d;
  int o = 0;
  A.oY = o; o += d.nspec; A.oYN = o; o += d.nvar; A.oF0 = o; o += d.nvar; A.oFC = o; o += d.nvar;
  A.oK = o; o += 6 * d.nvar; A.oGA = o; o += d.ng; A.oRCX = o; o += d.nrcx; A.oAB = o; o += d.ab_len;
  A.ws_stride = (size_t)o * 32;
  h->l_blocks = h->sm_count * bps;
  if (h->l_ws.ensure(sizeof(double) * A.ws_stride * h->l_blocks)) return fail(-1002, "out of device memory for the lane workspace (%zu MB)", (sizeof(double) * A.ws_stride * h->l_blocks) >> 20);
  A.ws = h->l_ws.as<double>();
  h->l_ready = 1;
  return 0;
}

static int choose_kernel_plain(gckpp_gpu_handle *h, const Decoded &d);
static int choose_kernel(gckpp_gpu_handle *h, const Decoded &d)
{
  const int k = choose_kernel_plain(h, d);
  // auto-reduce: the block kernel integrates on the full pattern with a per-cell keep mask (its AR instance); every
  // other choice falls back to the table-driven kernel, which carries the reference-order implementation
  if (d.autoreduce && k != 1) return 0;
  return k;
}
static int choose_kernel_plain(gckpp_gpu_handle *h, const Decoded &d)
{
  if (h->opt_kernel == 0) return 0;          // "kernel"=0 forces the table-driven, reference-order kernel
  if (h->T->nnz <= 0) return 0;
  if (h->opt_kernel == 3) return host_lsched(h->mech_id) ? 3 : 0;     // lane kernel: every method
  // small mechanisms (Hg: 32 species, 161 matrix entries): the lane kernel is the default -- a warp's whole
  // workspace is 150 KB, seven warps fit an SM, 2.7 M cells/s against 1.1 M for the block kernel (profiles/r02r_hg_bench.log)
  if (h->opt_kernel == 4) return unrolled_kernel_supports(h->mech_id) ? 4 : 0;
  if (h->opt_kernel < 0 && unrolled_kernel_supports(h->mech_id)) return 4;      // one cell per thread, straight-line code
  if (h->opt_kernel < 0 && h->T->nvar <= 64 && host_lsched(h->mech_id)) return 3;
  if (!(d.ICNTRL[2] == 0 || d.ICNTRL[2] == 4)) return 0;
  if (d.o.Tstart == d.o.Tend) return 0;
  // "kernel"=2: the warp-group kernel -- on explicit request only: its results vary from run to run at the 1e-9
  // level (37 % of the cells, tests/gpu_tools/determinism.py), the block-synchronous kernel's never do
  if (h->opt_kernel == 2) return (host_wsched(h->mech_id) && warp_kernel_supports(h->mech_id)) ? 2 : 0;
  return (host_sched(h->mech_id) && smem_kernel_supports(h->mech_id)) ? 1 : 0;
}

static int run_integrator(gckpp_gpu_handle *h, const Decoded &d, int ncell, int nwork, const int *cell_list,
                          const double *conc_in, const double *rconst, const double *hstart,
                          double *conc_out, int32_t *istatus, double *rstatus, int32_t *ierr,
                          int rc_stride = 0, int rc_cell0 = 0)
{
  if (nwork <= 0) return 0;
  int kern = choose_kernel(h, d);
  // the block kernel's auto-reduce instance closes with a pass that reads conc_in and ierr: without them, kernel 0
  if (d.autoreduce && kern == 1 && (!ierr || conc_in == conc_out)) kern = 0;
  const bool smem = kern == 1;
  int blocks = (nwork + h->threads - 1) / h->threads;
  if (blocks > h->max_blocks) blocks = h->max_blocks;
  int rc = kern == 4 ? 0 : kern == 3 ? prepare_lane(h) : kern == 2 ? prepare_warp(h) : (smem ? prepare_smem(h) : ensure_workspace(h, blocks));
  if (rc) return rc;
  RosArgs a;
  a.ncell = ncell; a.nwork = nwork; a.cell_list = cell_list;
  a.conc_in = conc_in; a.rconst = rconst; a.hstart = hstart;
  a.rc_stride = rc_stride > 0 ? rc_stride : ncell; a.rc_cell0 = rc_cell0;
  a.atol = h->tol.as<double>(); a.rtol = h->tol.as<double>() + h->T->nvar;
  a.conc_out = conc_out; a.istatus = istatus; a.rstatus = rstatus; a.ierr = ierr;
  a.work = h->work.as<double>(); a.ws_stride = (size_t)h->L.total * 32;
  a.next = h->next.as<int>(); a.sums = h->sums.as<unsigned long long>();
  a.L = h->L; a.o = d.o;
  // auto-reduce options as Rosenbrock() decodes them (gckpp_Integrator.F90:390-394, :479-486)
  a.ar_on = d.autoreduce; a.ar_target = d.ICNTRL[13]; a.ar_keep_active = h->keep_n > 0;
  a.ar_threshold = d.RCNTRL[11] > 0.0 ? d.RCNTRL[11] : 1.0e2; a.ar_ratio = d.RCNTRL[13];
  a.ar_keep_spc = h->keep_n > 0 ? h->keep_spc.as<unsigned char>() : nullptr;
  CUDA_TRY(cudaMemsetAsync(h->next.p, 0, sizeof(int), h->stream));
  if (h->T->nnz == 0) {   // carbon: forward Euler
    CUDA_TRY(launch_feuler(h->M, a, d.ICNTRL[15], h->stream));
  } else if (kern == 4) { // one cell per thread, persistent lanes
    const int bt = unrolled_block_threads();
    int nb = (nwork + bt - 1) / bt;
    if (nb > h->sm_count * unrolled_blocks_per_sm()) nb = h->sm_count * unrolled_blocks_per_sm();
    CUDA_TRY(launch_ros_unrolled(h->mech_id, a, nb, h->stream));
    h->last_kernel = 4;
  } else if (kern == 3) { // one cell per lane, one warp per block, persistent
    int nb = (nwork + 31) / 32;
    if (nb > h->l_blocks) nb = h->l_blocks;
    if (h->sm_blocks_cap > 0 && nb > h->sm_blocks_cap) nb = h->sm_blocks_cap;
    CUDA_TRY(launch_ros_lane(h->largs, a, nb, h->stream));
    h->last_kernel = 3;
  } else if (kern == 2) { // one persistent block per SM, one cell per warp
    const int cpb = warp_cells_per_block(h->mech_id);
    int nb = (nwork + cpb - 1) / cpb;
    if (nb > h->sm_count) nb = h->sm_count;
    if (h->sm_blocks_cap > 0 && nb > h->sm_blocks_cap) nb = h->sm_blocks_cap;
    CUDA_TRY(launch_ros_warp(h->mech_id, h->wargs, a, nb, h->stream));
    h->last_kernel = 2;
  } else if (smem) {      // one persistent block per SM, SMEM_NC cells each
    int nb = (nwork + SMEM_NC - 1) / SMEM_NC;
    if (nb > h->sm_count) nb = h->sm_count;
    if (h->sm_blocks_cap > 0 && nb > h->sm_blocks_cap) nb = h->sm_blocks_cap;
    if (d.autoreduce) {
      // the decision pass (Prod / LossY against the threshold at the initial state), then the AR instance; its two
      // extra pointers travel in RosArgs fields the block kernel does not read (see ros_smem.cu)
      if (h->ar_mask.ensure((size_t)h->T->nvar * (size_t)ncell)) return fail(-1002, "out of device memory for the auto-reduce masks");
      CUDA_TRY(launch_ar_mask(h->M, a, h->ar_mask.as<unsigned char>(), h->stream));
      a.work = reinterpret_cast<double *>(h->ar_mask.p);
      a.ar_keep_spc = reinterpret_cast<const unsigned char *>(h->sm_rowcol.p);
      h->stats[6] += 1;
    }
    CUDA_TRY(smem_set_directory(h->mech_id, h->plan.dir.data(), (int)h->plan.dir.size(), h->stream));     // per device, 250 bytes
    CUDA_TRY(launch_ros_smem(h->mech_id, h->sargs, a, nb, h->stream, d.autoreduce != 0));
    if (d.autoreduce) {
      CUDA_TRY(launch_ar_first_order(h->M, a, h->ar_mask.as<unsigned char>(), h->stream));
      h->stats[6] += 1;
    }
    h->last_kernel = 1;
  } else {
    h->last_kernel = 0;
    CUDA_TRY(launch_ros_generic(h->M, a, blocks, h->threads, h->stream));
  }
  h->stats[6] += 1;
  return 0;
}

// The arrays of one call (or one wave of a call) on the device; every per-cell array has row stride n.
struct DevIO {
  const double *conc_in, *rconst, *temp, *numden, *h2o, *photol, *khet, *hstart;
  const double *het;          // [GCKPP_NHET][n] HetState fields for the device-side heterogeneous laws, or NULL
  const uint8_t *active;
  double *conc_out; int32_t *istatus; double *rstatus; int32_t *ierr;
  double *rconst_work;        // [NREACT][n] scratch for Update_RCONST when rconst == NULL
};

static void print_profile(gckpp_gpu_handle *h, const unsigned long long *sums)
{
  if (h->last_kernel == 3) {
    fprintf(stderr, "[gckpp profile] lane kernel, block 0, cycles: refill %llu load_y %llu fun(x3) %llu jac_rates %llu jac_sums %llu lu %llu stage_vectors %llu solves(x4) %llu error+accept %llu\n",
            sums[8], sums[9], sums[10], sums[11], sums[12], sums[13], sums[14], sums[15], sums[16]);
  } else if (h->last_kernel == 2) {
    fprintf(stderr, "[gckpp profile] lead warp of group 0, block 0, cycles: load %llu fun0(vdot) %llu jac %llu lu_head %llu lu_tail %llu tail_solve %llu stage_rhs+vdot %llu solve_streams %llu accept %llu retire %llu rates %llu\n",
            sums[8], sums[9], sums[10], sums[11], sums[12], sums[13], sums[14], sums[15], sums[16], sums[17], sums[18]);
    const char *kn[4] = {"vdot", "jvs", "lu", "solve"};
    for (int k = 0; k < 4; k++)
      fprintf(stderr, "[gckpp profile]   %-5s bundles %llu: table wait %llu operands+fma %llu shuffles %llu store %llu barrier %llu\n", kn[k],
              sums[20 + 6 * k + 5], sums[20 + 6 * k + 0], sums[20 + 6 * k + 1], sums[20 + 6 * k + 2], sums[20 + 6 * k + 3], sums[20 + 6 * k + 4]);
  } else {
    fprintf(stderr, "[gckpp profile] block 0 cycles: control %llu fun(x3) %llu jac %llu lu_head %llu lu_tail %llu postlu %llu solve(rest) %llu accept %llu | solve: exec %llu prefetch %llu barrier %llu tails %llu | bundle: fetch+decode %llu terms %llu shuffles %llu write %llu\n",
            sums[8], sums[9], sums[10], sums[11], sums[12], sums[13], sums[14], sums[15], sums[16], sums[17], sums[18], sums[19],
            sums[20], sums[21], sums[22], sums[23]);
    fprintf(stderr, "[gckpp profile]   lu rounds:");
    for (int i = 0; i < 24; i++) fprintf(stderr, " %llu", sums[24 + i]);
    fprintf(stderr, "\n[gckpp profile]   sweep rounds (4 solves):");
    for (int i = 0; i < 24; i++) fprintf(stderr, " %llu", sums[48 + i]);
    fprintf(stderr, "\n");
  }
}

// One batch of n cells on the device: Update_RCONST (when no rate constants are given), the InChemGrid mask, the
// integration, and Do_FullChem's retry.  The counters in h->sums accumulate over the batches of a call (the caller
// clears them once); the handle stream is synchronised only when a count has to reach the host (mask, retry).
// stats[4] / stats[5] accumulate the retried / failed-twice cells.
static int device_core(gckpp_gpu_handle *h, const Decoded &d, int n, const DevIO &io, unsigned long long *fails_before)
{
  const gckpp_host_tables_t *T = h->T;
  if (!io.rconst && (!io.temp || !io.numden || !io.h2o)) return fail(-10, "rconst is NULL and temp/numden/h2o are not all given");
  // When the rate constants are computed here they live in a scratch of at most `device_wave_cells` columns
  // (8.5 KB per cell for fullchem): a batch larger than that is walked through in cell ranges.
  const int DW = io.rconst ? n : (h->opt_dev_wave > 0 ? h->opt_dev_wave : (1 << 20));
  const bool ranged = n > DW;
  if ((ranged || io.active || h->opt_retry) && h->cell_list.ensure(sizeof(int) * (size_t)n)) return fail(-1002, "out of device memory");
  if (ranged && !io.active) {
    if (h->ident.ensure(sizeof(int) * (size_t)n)) return fail(-1002, "out of device memory");
    if (h->ident_n < n) { CUDA_TRY(launch_iota(h->ident.as<int>(), n, h->stream)); h->ident_n = n; }
  }
  for (int c0 = 0; c0 < n; c0 += DW) {
    const int m = (c0 + DW <= n) ? DW : n - c0, c1 = c0 + m;
    const double *rconst = io.rconst;
    int rc_stride = n, rc_cell0 = 0;
    if (!rconst) {
      CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
      const bool dohet = io.het && h->srmw_n > 0;
      CUDA_TRY(launch_update_rconst(h->mech_id, m, io.temp + c0, io.numden + c0, io.h2o + c0, io.photol ? io.photol + c0 : nullptr,
                                    io.khet ? io.khet + c0 : nullptr, io.rconst_work, h->stream, /*input stride*/ n, /*output stride*/ m,
                                    dohet ? io.het + c0 : nullptr, dohet ? io.conc_in + c0 : nullptr, dohet ? h->srmw.as<double>() : nullptr,
                                    h->spc_data_n));
      CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
      h->stats[6] += 1;
      h->stats[12] += 1;          // Update_RCONST launches of this call
      rconst = io.rconst_work; rc_stride = m; rc_cell0 = c0;
    }
    // cells outside the chemistry grid: copy through, zero status
    const int *cell_list = nullptr;
    int nwork = m;
    if (io.active) {
      CUDA_TRY(cudaMemsetAsync(h->counter.p, 0, sizeof(int), h->stream));
      CUDA_TRY(launch_select_active(n, c0, c1, io.active, T->nspec, io.conc_in, io.conc_out, io.istatus, io.rstatus, io.ierr,
                                    h->cell_list.as<int>(), h->counter.as<int>(), h->stream));
      h->stats[6] += 1;
      CUDA_TRY(cudaMemcpyAsync(&nwork, h->counter.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      cell_list = h->cell_list.as<int>();
    } else if (ranged) {
      cell_list = h->ident.as<int>() + c0;
    }
    int rc = run_integrator(h, d, n, nwork, cell_list, io.conc_in, rconst, io.hstart, io.conc_out, io.istatus, io.rstatus, io.ierr,
                            rc_stride, rc_cell0);
    if (rc) return rc;
    // Do_FullChem's retry (fullchem_mod.F90:1138-1162): C restored, RCNTRL(3) = 0, integrate again.  (The reference
    // also sets RCNTRL(12) = -1 to switch auto-reduce off, but Integrate's merge drops non-positive RCNTRL values
    // (gckpp_Integrator.F90:116), so auto-reduce stays on for the retry there too: SURVEY Q3.)
    if (h->opt_retry && io.ierr) {
      unsigned long long sums[4];
      CUDA_TRY(cudaMemcpyAsync(sums, h->sums.p, sizeof sums, cudaMemcpyDeviceToHost, h->stream));
      CUDA_TRY(cudaStreamSynchronize(h->stream));
      const int nfail = (int)(sums[2] - *fails_before);
      if (nfail > 0) {
        CUDA_TRY(cudaMemsetAsync(h->counter.p, 0, sizeof(int), h->stream));
        CUDA_TRY(launch_select_failed(c0, c1, io.ierr, h->cell_list.as<int>(), h->counter.as<int>(), h->stream));
        int nretry = 0;
        CUDA_TRY(cudaMemcpyAsync(&nretry, h->counter.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        Decoded d2 = d;
        d2.o.Hstart_rcntrl = 0.0;
        rc = run_integrator(h, d2, n, nretry, h->cell_list.as<int>(), io.conc_in, rconst, nullptr, io.conc_out, io.istatus,
                            io.rstatus, io.ierr, rc_stride, rc_cell0);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(sums, h->sums.p, sizeof sums, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->stats[4] += nretry;
        h->stats[5] += (double)(sums[2] - *fails_before - (unsigned long long)nfail);    // failed again
        h->stats[6] += 1;
      }
      *fails_before = sums[2];
    }
  }
  return 0;
}

// option checks shared by the entry points; fills ierr on the reference's own option errors (-1..-5)
static int check_decoded(gckpp_gpu_handle *h, const Decoded &d)
{
  if (d.autoreduce && d.ICNTRL[12] == 1) return fail(-12, "the append variant of auto-reduce (ICNTRL(13)=1) is not available in this build");
  if (d.autoreduce && !h->T->fun_split) return fail(-12, "auto-reduce needs the split ODE function (fullchem)");
  return 0;
}

static int finish_stats(gckpp_gpu_handle *h)
{
  unsigned long long sums[128];
  CUDA_TRY(cudaMemcpyAsync(sums, h->sums.p, sizeof sums, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  if (getenv("GCKPP_PROFILE")) print_profile(h, sums);
  h->stats[3] = (double)sums[3]; h->stats[7] = (double)sums[0]; h->stats[8] = (double)sums[1];
  h->stats[10] = (double)sums[2];           // integrations that ended with IERR < 0 (first pass and retry)
  h->stats[13] = h->last_kernel;
  if (h->stats[12] > 0) {                   // Update_RCONST: the last launch's time, scaled to the launches of the call
    float ms = 0;
    if (cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]) == cudaSuccess) h->stats[1] = ms * h->stats[12];
    else cudaGetLastError();
  }
  return 0;
}

extern "C" int gckpp_gpu_integrate_device(gckpp_gpu_handle_t *h, int ncell, double tin, double tout,
                                          const double *conc_in, const double *rconst,
                                          const double *temp, const double *numden, const double *h2o,
                                          const double *photol, const double *khet,
                                          const double *atol, const double *rtol,
                                          const int32_t *icntrl, const double *rcntrl,
                                          const double *hstart, const uint8_t *active,
                                          double *conc_out, int32_t *istatus, double *rstatus, int32_t *ierr)
{
  if (!h) return fail(-10, "NULL handle");
  if (ncell < 0 || !conc_in || !conc_out) return fail(-10, "gckpp_gpu_integrate: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const gckpp_host_tables_t *T = h->T;
  Decoded d;
  int rc = decode_options(T, tin, tout, icntrl, rcntrl, atol, rtol, d);
  if (rc) {
    if (ierr && rc >= -5) CUDA_TRY(launch_fill_int(ierr, ncell, rc, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return rc;
  }
  if ((rc = check_decoded(h, d))) return rc;
  for (int i = 0; i < 16; i++) h->stats[i] = 0.0;
  if (atol && rtol) {
    CUDA_TRY(cudaMemcpyAsync(h->tol.p, atol, sizeof(double) * T->nvar, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->tol.as<double>() + T->nvar, rtol, sizeof(double) * T->nvar, cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_TRY(cudaMemsetAsync(h->sums.p, 0, 128 * sizeof(unsigned long long), h->stream));
  {
    const size_t dw = (size_t)(h->opt_dev_wave > 0 ? h->opt_dev_wave : (1 << 20));
    if (!rconst && h->rconst_work.ensure(sizeof(double) * (size_t)T->nreact * ((size_t)ncell < dw ? (size_t)ncell : dw))) return fail(-1002, "out of device memory for rconst");
  }
  DevIO io{conc_in, rconst, temp, numden, h2o, photol, khet, hstart, h->het_user, active, conc_out, istatus, rstatus, ierr,
           h->rconst_work.as<double>()};
  CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
  unsigned long long fails = 0;
  if ((rc = device_core(h, d, ncell, io, &fails))) return rc;
  CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));
  if ((rc = finish_stats(h))) return rc;
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[0], h->ev[3]);
  h->stats[0] = ms; h->stats[9] = ms;
  return h->opt_retry ? (int)h->stats[5] : 0;
}

// ---- host-buffer entry ------------------------------------------------------------------------------------
// The cells are processed in WAVES of at most `wave_cells` cells (contiguous cell ranges; every per-cell array is
// [rows][ncell] cell-fastest, so a wave is a 2-D copy with pitch ncell).  Device memory is bounded by two waves, whatever
// the grid (C180: 14 M cells), and the copies overlap the integration: while wave i integrates on the handle stream,
// the inputs of wave i+1 travel to the device on the copy-in stream and the results of wave i-1 travel back on the
// copy-out stream.  This is the batched replacement of the cell loop of Do_FullChem (fullchem_mod.F90:528-1551): one
// call per chemistry step per GPU, as GCHP calls the routine once per tile (gchp_chunk_mod.F90:1366).
struct WaveSlot {
  DevBuf conc_in, conc_out, rconst, met, photol, khet, het, hstart, active, ist, rst, ierr;
  cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
};

static int ensure_slots(gckpp_gpu_handle *h)
{
  if (h->slots) return 0;
  h->slots = new WaveSlot[2];
  for (int i = 0; i < 2; i++) {
    CUDA_TRY(cudaEventCreateWithFlags(&h->slots[i].ev_in, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->slots[i].ev_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->slots[i].ev_out, cudaEventDisableTiming));
  }
  if (!h->s_in) CUDA_TRY(cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking));
  if (!h->s_out) CUDA_TRY(cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking));
  return 0;
}

static void free_slots(gckpp_gpu_handle *h)
{
  if (!h->slots) return;
  for (int i = 0; i < 2; i++) {
    WaveSlot &s = h->slots[i];
    DevBuf *b[] = {&s.conc_in, &s.conc_out, &s.rconst, &s.met, &s.photol, &s.khet, &s.het, &s.hstart, &s.active, &s.ist, &s.rst, &s.ierr};
    for (DevBuf *x : b) x->release();
    if (s.ev_in) cudaEventDestroy(s.ev_in);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_out) cudaEventDestroy(s.ev_out);
  }
  delete[] h->slots;
  h->slots = nullptr;
}

// "pin"=1: page-lock the caller's arrays for the duration of the call (Fortran State_Chm arrays are pageable; copies
// from pageable memory are staged by the driver and do not overlap).  Registration failures are not errors.
struct HostPins {
  std::vector<void *> p;
  void add(const void *ptr, size_t bytes)
  {
    if (ptr && bytes && cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess) p.push_back(const_cast<void *>(ptr));
    else cudaGetLastError();
  }
  ~HostPins() { for (void *q : p) cudaHostUnregister(q); }
};

extern "C" int gckpp_gpu_integrate(gckpp_gpu_handle_t *h, int ncell, double tin, double tout,
                                   const double *conc_in, const double *rconst,
                                   const double *temp, const double *numden, const double *h2o,
                                   const double *photol, const double *khet,
                                   const double *atol, const double *rtol,
                                   const int32_t *icntrl, const double *rcntrl,
                                   const double *hstart, const uint8_t *active,
                                   double *conc_out, int32_t *istatus, double *rstatus, int32_t *ierr)
{
  if (!h) return fail(-10, "NULL handle");
  if (ncell < 0 || !conc_in || !conc_out) return fail(-10, "gckpp_gpu_integrate: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const gckpp_host_tables_t *T = h->T;
  const size_t nc = (size_t)ncell;
  Decoded d;
  int rc = decode_options(T, tin, tout, icntrl, rcntrl, atol, rtol, d);
  if (rc) {
    if (ierr && rc >= -5) for (size_t i = 0; i < nc; i++) ierr[i] = rc;
    return rc;
  }
  if ((rc = check_decoded(h, d))) return rc;
  if (!rconst && (!temp || !numden || !h2o)) return fail(-10, "rconst is NULL and temp/numden/h2o are not all given");
  if ((rc = ensure_slots(h))) return rc;
  for (int i = 0; i < 16; i++) h->stats[i] = 0.0;
  const bool have_met = temp && numden && h2o, have_ph = photol && T->nphot, have_kh = khet && T->next;
  const double *het = (h->srmw_n > 0 && !rconst) ? h->het_user : nullptr;      // HetState fields (host array) for the device-side laws
  // wave size: "wave_cells" (default 65536), or ncell / "chunks" when that option was set explicitly
  size_t W = (size_t)(h->opt_wave_cells > 0 ? h->opt_wave_cells : 65536);
  if (h->opt_chunks > 0) W = (nc + h->opt_chunks - 1) / h->opt_chunks;
  if (W > nc) W = nc;
  if (W < 1) W = 1;
  const int nwaves = (int)((nc + W - 1) / W);
  HostPins pins;
  if (h->opt_pin) {
    pins.add(conc_in, 8 * T->nspec * nc); pins.add(conc_out, 8 * T->nspec * nc);
    if (rconst) pins.add(rconst, 8 * T->nreact * nc);
    if (have_met) { pins.add(temp, 8 * nc); pins.add(numden, 8 * nc); pins.add(h2o, 8 * nc); }
    if (have_ph) pins.add(photol, 8 * T->nphot * nc);
    if (have_kh) pins.add(khet, 8 * T->next * nc);
    if (hstart) pins.add(hstart, 8 * nc);
    if (het) pins.add(het, 8 * (size_t)GCKPP_NHET * nc);
    if (istatus) pins.add(istatus, 4 * 8 * nc);
    if (rstatus) pins.add(rstatus, 8 * 4 * nc);
    if (ierr) pins.add(ierr, 4 * nc);
  }
  for (int i = 0; i < 2 && i < nwaves; i++) {
    WaveSlot &s = h->slots[i];
    if (s.conc_in.ensure(8 * T->nspec * W) || s.conc_out.ensure(8 * T->nspec * W) || s.ist.ensure(4 * 8 * W) ||
        s.rst.ensure(8 * 4 * W) || s.ierr.ensure(4 * W) || s.rconst.ensure(8 * T->nreact * W) ||
        (have_met && s.met.ensure(3 * 8 * W)) || (hstart && s.hstart.ensure(8 * W)) || (active && s.active.ensure(W)) ||
        (have_ph && s.photol.ensure(8 * T->nphot * W)) || (have_kh && s.khet.ensure(8 * T->next * W)) ||
        (het && s.het.ensure(8 * (size_t)GCKPP_NHET * W)))
      return fail(-1002, "out of device memory for a wave of %zu cells", W);
  }
  CUDA_TRY(cudaMemcpyAsync(h->tol.p, atol, sizeof(double) * T->nvar, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(h->tol.as<double>() + T->nvar, rtol, sizeof(double) * T->nvar, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemsetAsync(h->sums.p, 0, 128 * sizeof(unsigned long long), h->stream));
  // rows x n elements of a [rows][ncell] host array <-> a [rows][n] device array
  auto rows_in = [&](void *dst, const void *src, size_t c0, size_t n, size_t rows, size_t elt) {
    return cudaMemcpy2DAsync(dst, n * elt, (const char *)src + c0 * elt, nc * elt, n * elt, rows, cudaMemcpyHostToDevice, h->s_in);
  };
  auto rows_out = [&](void *dst, const void *src, size_t c0, size_t n, size_t rows, size_t elt) {
    return cudaMemcpy2DAsync((char *)dst + c0 * elt, nc * elt, src, n * elt, n * elt, rows, cudaMemcpyDeviceToHost, h->s_out);
  };
  auto copy_in = [&](int w) -> int {
    WaveSlot &s = h->slots[w & 1];
    const size_t c0 = (size_t)w * W, n = (c0 + W <= nc) ? W : nc - c0;
    if (w >= 2) CUDA_TRY(cudaStreamWaitEvent(h->s_in, s.ev_out, 0));      // the slot's previous results have left
    CUDA_TRY(rows_in(s.conc_in.p, conc_in, c0, n, T->nspec, 8));
    if (rconst) CUDA_TRY(rows_in(s.rconst.p, rconst, c0, n, T->nreact, 8));
    if (have_met) {
      double *m = s.met.as<double>();
      CUDA_TRY(rows_in(m, temp, c0, n, 1, 8)); CUDA_TRY(rows_in(m + n, numden, c0, n, 1, 8)); CUDA_TRY(rows_in(m + 2 * n, h2o, c0, n, 1, 8));
    }
    if (have_ph) CUDA_TRY(rows_in(s.photol.p, photol, c0, n, T->nphot, 8));
    if (have_kh) CUDA_TRY(rows_in(s.khet.p, khet, c0, n, T->next, 8));
    if (het) CUDA_TRY(rows_in(s.het.p, het, c0, n, GCKPP_NHET, 8));
    if (hstart) CUDA_TRY(rows_in(s.hstart.p, hstart, c0, n, 1, 8));
    if (active) CUDA_TRY(rows_in(s.active.p, active, c0, n, 1, 1));
    CUDA_TRY(cudaEventRecord(s.ev_in, h->s_in));
    return 0;
  };
  CUDA_TRY(cudaEventRecord(h->ev[4], h->stream));
  if ((rc = copy_in(0))) return rc;
  unsigned long long fails = 0;
  for (int w = 0; w < nwaves; w++) {
    WaveSlot &s = h->slots[w & 1];
    const size_t c0 = (size_t)w * W, n = (c0 + W <= nc) ? W : nc - c0;
    if (w + 1 < nwaves && (rc = copy_in(w + 1))) return rc;               // runs ahead of the integration of wave w
    CUDA_TRY(cudaStreamWaitEvent(h->stream, s.ev_in, 0));
    double *m = s.met.as<double>();
    DevIO io{s.conc_in.as<double>(), rconst ? s.rconst.as<double>() : nullptr, have_met ? m : nullptr, have_met ? m + n : nullptr,
             have_met ? m + 2 * n : nullptr, have_ph ? s.photol.as<double>() : nullptr, have_kh ? s.khet.as<double>() : nullptr,
             hstart ? s.hstart.as<double>() : nullptr, het ? s.het.as<double>() : nullptr,
             active ? s.active.as<uint8_t>() : nullptr, s.conc_out.as<double>(),
             s.ist.as<int32_t>(), s.rst.as<double>(), s.ierr.as<int32_t>(), s.rconst.as<double>()};
    if ((rc = device_core(h, d, (int)n, io, &fails))) return rc;
    CUDA_TRY(cudaEventRecord(s.ev_done, h->stream));
    CUDA_TRY(cudaStreamWaitEvent(h->s_out, s.ev_done, 0));
    CUDA_TRY(rows_out(conc_out, s.conc_out.p, c0, n, T->nspec, 8));
    if (istatus) CUDA_TRY(rows_out(istatus, s.ist.p, c0, n, 8, 4));
    if (rstatus) CUDA_TRY(rows_out(rstatus, s.rst.p, c0, n, 4, 8));
    if (ierr) CUDA_TRY(rows_out(ierr, s.ierr.p, c0, n, 1, 4));
    CUDA_TRY(cudaEventRecord(s.ev_out, h->s_out));
  }
  CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));
  if ((rc = finish_stats(h))) return rc;
  CUDA_TRY(cudaStreamSynchronize(h->s_out));
  float ms = 0;
  cudaEventElapsedTime(&ms, h->ev[4], h->ev[3]);
  h->stats[0] = ms; h->stats[9] = ms; h->stats[11] = nwaves;
  return h->opt_retry ? (int)h->stats[5] : 0;
}

static int h2d(gckpp_gpu_handle *h, DevBuf &b, const void *src, size_t bytes)
{
  if (b.ensure(bytes)) return fail(-1002, "out of device memory (%zu MB)", bytes >> 20);
  CUDA_TRY(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return 0;
}

// ---- heterogeneous laws on the device ---------------------------------------------------------------------
extern "C" int gckpp_gpu_set_sr_mw(gckpp_gpu_handle_t *h, int n, const double *sr_mw)
{
  if (!h || n < 0 || (n > 0 && !sr_mw)) return fail(-10, "gckpp_gpu_set_sr_mw: bad arguments");
  if (n != 0 && n != h->T->nspec) return fail(-10, "gckpp_gpu_set_sr_mw: expected %d values (one per species)", h->T->nspec);
  CUDA_TRY(cudaSetDevice(h->device));
  if (n > 0) {
    if (h->srmw.ensure(sizeof(double) * 4 * (size_t)n)) return fail(-1002, "out of device memory");
    CUDA_TRY(cudaMemcpy(h->srmw.p, sr_mw, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  }
  h->srmw_n = n;
  h->spc_data_n = 0;
  return 0;
}

extern "C" int gckpp_gpu_set_species_data(gckpp_gpu_handle_t *h, int n, const double *sr_mw, const double *mw,
                                          const double *henry_k0, const double *henry_cr)
{
  if (!h || !sr_mw || !mw || !henry_k0 || !henry_cr) return fail(-10, "gckpp_gpu_set_species_data: bad arguments");
  int rc = gckpp_gpu_set_sr_mw(h, n, sr_mw);
  if (rc || n == 0) return rc;
  const double *src[3] = {mw, henry_k0, henry_cr};
  for (int k = 0; k < 3; k++)
    CUDA_TRY(cudaMemcpy(h->srmw.as<double>() + (size_t)(k + 1) * n, src[k], sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
  h->spc_data_n = n;
  return 0;
}

extern "C" int gckpp_gpu_set_het(gckpp_gpu_handle_t *h, const double *het, const double *conc)
{
  if (!h) return fail(-10, "NULL handle");
  if (het && h->mech_id != GCKPP_MECH_FULLCHEM) return fail(-11, "device-side heterogeneous laws exist for fullchem only");
  if (het && h->srmw_n == 0) return fail(-10, "gckpp_gpu_set_het: call gckpp_gpu_set_sr_mw first");
  h->het_user = het; h->het_conc_user = het ? conc : nullptr;
  return 0;
}

extern "C" int gckpp_gpu_update_rconst_device(gckpp_gpu_handle_t *h, int ncell,
                                              const double *temp, const double *numden, const double *h2o,
                                              const double *photol, const double *khet, double *rconst_out)
{
  if (!h || !temp || !numden || !h2o || !rconst_out || ncell < 0) return fail(-10, "gckpp_gpu_update_rconst: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const bool dohet = h->het_user && h->het_conc_user && h->srmw_n > 0;        // both device pointers here
  CUDA_TRY(launch_update_rconst(h->mech_id, ncell, temp, numden, h2o, photol, khet, rconst_out, h->stream, 0, 0,
                                dohet ? h->het_user : nullptr, dohet ? h->het_conc_user : nullptr, dohet ? h->srmw.as<double>() : nullptr,
                                h->spc_data_n));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_update_rconst(gckpp_gpu_handle_t *h, int ncell,
                                       const double *temp, const double *numden, const double *h2o,
                                       const double *photol, const double *khet, double *rconst_out)
{
  if (!h || !temp || !numden || !h2o || !rconst_out || ncell < 0) return fail(-10, "gckpp_gpu_update_rconst: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const gckpp_host_tables_t *T = h->T;
  const size_t nc = (size_t)ncell;
  int rc;
  if (h->s_met.ensure(3 * sizeof(double) * nc)) return fail(-1002, "out of device memory");
  double *d_temp = h->s_met.as<double>(), *d_numden = d_temp + nc, *d_h2o = d_numden + nc;
  CUDA_TRY(cudaMemcpyAsync(d_temp, temp, sizeof(double) * nc, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(d_numden, numden, sizeof(double) * nc, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(d_h2o, h2o, sizeof(double) * nc, cudaMemcpyHostToDevice, h->stream));
  if (photol && T->nphot && (rc = h2d(h, h->s_photol, photol, sizeof(double) * T->nphot * nc))) return rc;
  if (khet && T->next && (rc = h2d(h, h->s_khet, khet, sizeof(double) * T->next * nc))) return rc;
  if (h->s_rconst.ensure(sizeof(double) * T->nreact * nc)) return fail(-1002, "out of device memory");
  const bool dohet = h->het_user && h->het_conc_user && h->srmw_n > 0;        // both host arrays here: staged
  if (dohet && ((rc = h2d(h, h->s_conc_in, h->het_conc_user, sizeof(double) * T->nspec * nc)) ||
                (rc = h2d(h, h->s_conc_out, h->het_user, sizeof(double) * (size_t)GCKPP_NHET * nc)))) return rc;
  CUDA_TRY(launch_update_rconst(h->mech_id, ncell, d_temp, d_numden, d_h2o,
                                (photol && T->nphot) ? h->s_photol.as<double>() : nullptr,
                                (khet && T->next) ? h->s_khet.as<double>() : nullptr, h->s_rconst.as<double>(), h->stream, 0, 0,
                                dohet ? h->s_conc_out.as<double>() : nullptr, dohet ? h->s_conc_in.as<double>() : nullptr,
                                dohet ? h->srmw.as<double>() : nullptr, h->spc_data_n));
  CUDA_TRY(cudaMemcpyAsync(rconst_out, h->s_rconst.p, sizeof(double) * T->nreact * nc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_fun(gckpp_gpu_handle_t *h, int ncell, const double *conc, const double *rconst,
                             double *vdot, double *aout)
{
  if (!h || !conc || !rconst || ncell < 0) return fail(-10, "gckpp_gpu_fun: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const gckpp_host_tables_t *T = h->T;
  const size_t nc = (size_t)ncell;
  int rc;
  if ((rc = h2d(h, h->s_conc_in, conc, sizeof(double) * T->nspec * nc))) return rc;
  if ((rc = h2d(h, h->s_rconst, rconst, sizeof(double) * T->nreact * nc))) return rc;
  if (h->s_conc_out.ensure(sizeof(double) * T->nspec * nc) || h->scratch.ensure(sizeof(double) * T->nreact * nc))
    return fail(-1002, "out of device memory");
  CUDA_TRY(launch_fun_cells(h->M, ncell, h->s_conc_in.as<double>(), h->s_rconst.as<double>(),
                            h->s_conc_out.as<double>(), h->scratch.as<double>(), h->stream));
  if (vdot) CUDA_TRY(cudaMemcpyAsync(vdot, h->s_conc_out.p, sizeof(double) * T->nvar * nc, cudaMemcpyDeviceToHost, h->stream));
  if (aout) CUDA_TRY(cudaMemcpyAsync(aout, h->scratch.p, sizeof(double) * T->nreact * nc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

// ---- the pieces of Do_FullChem around the integration (post.cu) -------------------------------------------------
// the small host lists of a call, packed into one device buffer: [ints][doubles][bytes]
static int stage_small(gckpp_gpu_handle *h, const int32_t *ints, int ni, const double *dbl, int nd, const uint8_t *bytes, int nb,
                       const int **d_int, const double **d_dbl, const unsigned char **d_bytes)
{
  const size_t oi = 0, od = ((size_t)ni * 4 + 7) & ~(size_t)7, ob = od + (size_t)nd * 8, tot = ob + (size_t)nb + 16;
  if (h->small.ensure(tot)) return fail(-1002, "out of device memory");
  char *base = (char *)h->small.p;
  if (ni) CUDA_TRY(cudaMemcpyAsync(base + oi, ints, (size_t)ni * 4, cudaMemcpyHostToDevice, h->stream));
  if (nd) CUDA_TRY(cudaMemcpyAsync(base + od, dbl, (size_t)nd * 8, cudaMemcpyHostToDevice, h->stream));
  if (nb) CUDA_TRY(cudaMemcpyAsync(base + ob, bytes, (size_t)nb, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));       // the host lists may be temporaries of the caller
  *d_int = (const int *)(base + oi); *d_dbl = (const double *)(base + od); *d_bytes = nb ? (const unsigned char *)(base + ob) : nullptr;
  return 0;
}
static int check_ids(const gckpp_host_tables_t *T, const int32_t *ids, int n, const char *what)
{
  if (n < 0 || (n > 0 && !ids)) return fail(-10, "%s: bad index list", what);
  for (int k = 0; k < n; k++) if (ids[k] < 0 || ids[k] >= T->nspec) return fail(-10, "%s: species index %d out of range", what, ids[k]);
  return 0;
}

extern "C" int gckpp_gpu_zero_species_device(gckpp_gpu_handle_t *h, int ncell, double *conc, int n, const int32_t *ids0)
{
  if (!h || !conc || ncell < 0) return fail(-10, "gckpp_gpu_zero_species: bad arguments");
  int rc;
  if ((rc = check_ids(h->T, ids0, n, "gckpp_gpu_zero_species"))) return rc;
  if (ncell == 0 || n == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const int *di; const double *dd; const unsigned char *db;
  if ((rc = stage_small(h, ids0, n, nullptr, 0, nullptr, 0, &di, &dd, &db))) return rc;
  CUDA_TRY(launch_zero_species(conc, ncell, di, n, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_post_integrate_device(gckpp_gpu_handle_t *h, int ncell, double *conc, int nscale, const int32_t *scale_ids0,
                                               const double *scale_div, const uint8_t *spc_mask, float *negatives)
{
  if (!h || !conc || ncell < 0 || (nscale > 0 && !scale_div)) return fail(-10, "gckpp_gpu_post_integrate: bad arguments");
  int rc;
  if ((rc = check_ids(h->T, scale_ids0, nscale, "gckpp_gpu_post_integrate"))) return rc;
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const int *di; const double *dd; const unsigned char *db;
  if ((rc = stage_small(h, scale_ids0, nscale, scale_div, nscale, spc_mask, spc_mask ? h->T->nspec : 0, &di, &dd, &db))) return rc;
  CUDA_TRY(launch_post_integrate(conc, ncell, h->T->nspec, di, dd, nscale, db, negatives, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_prod_loss_device(gckpp_gpu_handle_t *h, int ncell, const double *conc, double dt, int nslots,
                                          const int32_t *ids0, double *out)
{
  if (!h || !conc || !out || ncell < 0 || !(dt > 0.0)) return fail(-10, "gckpp_gpu_prod_loss: bad arguments");
  int rc;
  if ((rc = check_ids(h->T, ids0, nslots, "gckpp_gpu_prod_loss"))) return rc;
  if (ncell == 0 || nslots == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const int *di; const double *dd; const unsigned char *db;
  if ((rc = stage_small(h, ids0, nslots, nullptr, 0, nullptr, 0, &di, &dd, &db))) return rc;
  CUDA_TRY(launch_prod_loss(conc, ncell, dt, di, nslots, out, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_oh_reactivity_device(gckpp_gpu_handle_t *h, int ncell, const double *conc, const double *rconst, double *ohreact)
{
  if (!h || !conc || !rconst || !ohreact || ncell < 0) return fail(-10, "gckpp_gpu_oh_reactivity: bad arguments");
  if (h->T->nohr <= 0) return fail(-11, "the mechanism has no Get_OHreactivity");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(launch_oh_reactivity(conc, rconst, ncell, h->ohr_coef, h->ohr_rxn, h->ohr_spc, h->T->nohr, ohreact, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

// host-array variants: stage conc (and rconst) through the handle's buffers
extern "C" int gckpp_gpu_zero_species(gckpp_gpu_handle_t *h, int ncell, double *conc, int n, const int32_t *ids0)
{
  if (!h || !conc || ncell < 0) return fail(-10, "gckpp_gpu_zero_species: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * h->T->nspec * (size_t)ncell;
  int rc;
  if ((rc = h2d(h, h->s_conc_in, conc, bytes))) return rc;
  if ((rc = gckpp_gpu_zero_species_device(h, ncell, h->s_conc_in.as<double>(), n, ids0))) return rc;
  CUDA_TRY(cudaMemcpy(conc, h->s_conc_in.p, bytes, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int gckpp_gpu_post_integrate(gckpp_gpu_handle_t *h, int ncell, double *conc, int nscale, const int32_t *scale_ids0,
                                        const double *scale_div, const uint8_t *spc_mask, float *negatives)
{
  if (!h || !conc || ncell < 0) return fail(-10, "gckpp_gpu_post_integrate: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const size_t bytes = sizeof(double) * h->T->nspec * (size_t)ncell;
  int rc;
  if ((rc = h2d(h, h->s_conc_in, conc, bytes))) return rc;
  if (negatives && (rc = h2d(h, h->s_hstart, negatives, sizeof(float) * (size_t)ncell))) return rc;
  if ((rc = gckpp_gpu_post_integrate_device(h, ncell, h->s_conc_in.as<double>(), nscale, scale_ids0, scale_div, spc_mask,
                                            negatives ? h->s_hstart.as<float>() : nullptr))) return rc;
  CUDA_TRY(cudaMemcpy(conc, h->s_conc_in.p, bytes, cudaMemcpyDeviceToHost));
  if (negatives) CUDA_TRY(cudaMemcpy(negatives, h->s_hstart.p, sizeof(float) * (size_t)ncell, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int gckpp_gpu_prod_loss(gckpp_gpu_handle_t *h, int ncell, const double *conc, double dt, int nslots,
                                   const int32_t *ids0, double *out)
{
  if (!h || !conc || !out || ncell < 0 || nslots < 0) return fail(-10, "gckpp_gpu_prod_loss: bad arguments");
  if (ncell == 0 || nslots == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = h2d(h, h->s_conc_in, conc, sizeof(double) * h->T->nspec * (size_t)ncell))) return rc;
  if (h->s_conc_out.ensure(sizeof(double) * (size_t)nslots * ncell)) return fail(-1002, "out of device memory");
  if ((rc = gckpp_gpu_prod_loss_device(h, ncell, h->s_conc_in.as<double>(), dt, nslots, ids0, h->s_conc_out.as<double>()))) return rc;
  CUDA_TRY(cudaMemcpy(out, h->s_conc_out.p, sizeof(double) * (size_t)nslots * ncell, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int gckpp_gpu_oh_reactivity(gckpp_gpu_handle_t *h, int ncell, const double *conc, const double *rconst, double *ohreact)
{
  if (!h || !conc || !rconst || !ohreact || ncell < 0) return fail(-10, "gckpp_gpu_oh_reactivity: bad arguments");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  int rc;
  if ((rc = h2d(h, h->s_conc_in, conc, sizeof(double) * h->T->nspec * (size_t)ncell))) return rc;
  if ((rc = h2d(h, h->s_rconst, rconst, sizeof(double) * h->T->nreact * (size_t)ncell))) return rc;
  if (h->s_rst.ensure(sizeof(double) * (size_t)ncell)) return fail(-1002, "out of device memory");
  if ((rc = gckpp_gpu_oh_reactivity_device(h, ncell, h->s_conc_in.as<double>(), h->s_rconst.as<double>(), h->s_rst.as<double>()))) return rc;
  CUDA_TRY(cudaMemcpy(ohreact, h->s_rst.p, sizeof(double) * (size_t)ncell, cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int gckpp_gpu_jac(gckpp_gpu_handle_t *h, int ncell, const double *conc, const double *rconst, double *jvs)
{
  if (!h || !conc || !rconst || !jvs || ncell < 0) return fail(-10, "gckpp_gpu_jac: bad arguments");
  if (h->T->nnz == 0) return fail(-11, "mechanism has no Jacobian");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const gckpp_host_tables_t *T = h->T;
  const size_t nc = (size_t)ncell;
  int rc;
  if ((rc = h2d(h, h->s_conc_in, conc, sizeof(double) * T->nspec * nc))) return rc;
  if ((rc = h2d(h, h->s_rconst, rconst, sizeof(double) * T->nreact * nc))) return rc;
  if (h->scratch.ensure(sizeof(double) * ((size_t)T->nb + T->nnz) * nc)) return fail(-1002, "out of device memory");
  double *B = h->scratch.as<double>(), *J = B + (size_t)T->nb * nc;
  CUDA_TRY(launch_jac_cells(h->M, ncell, h->s_conc_in.as<double>(), h->s_rconst.as<double>(), B, J, h->stream));
  CUDA_TRY(cudaMemcpyAsync(jvs, J, sizeof(double) * T->nnz * nc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_decomp(gckpp_gpu_handle_t *h, int ncell, double *jvs, int32_t *ier)
{
  if (!h || !jvs || !ier || ncell < 0) return fail(-10, "gckpp_gpu_decomp: bad arguments");
  if (h->T->nnz == 0) return fail(-11, "mechanism has no Jacobian");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const gckpp_host_tables_t *T = h->T;
  const size_t nc = (size_t)ncell;
  if (h->scratch.ensure(sizeof(double) * ((size_t)T->nvar + T->nnz) * nc + sizeof(int) * nc)) return fail(-1002, "out of device memory");
  double *J = h->scratch.as<double>(), *W = J + (size_t)T->nnz * nc;
  int *E = (int *)(W + (size_t)T->nvar * nc);
  CUDA_TRY(cudaMemcpyAsync(J, jvs, sizeof(double) * T->nnz * nc, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(launch_decomp_cells(h->M, ncell, J, W, E, h->stream));
  CUDA_TRY(cudaMemcpyAsync(jvs, J, sizeof(double) * T->nnz * nc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaMemcpyAsync(ier, E, sizeof(int) * nc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_solve(gckpp_gpu_handle_t *h, int ncell, const double *jvs, double *x)
{
  if (!h || !jvs || !x || ncell < 0) return fail(-10, "gckpp_gpu_solve: bad arguments");
  if (h->T->nnz == 0) return fail(-11, "mechanism has no Jacobian");
  if (ncell == 0) return 0;
  CUDA_TRY(cudaSetDevice(h->device));
  const gckpp_host_tables_t *T = h->T;
  const size_t nc = (size_t)ncell;
  if (h->scratch.ensure(sizeof(double) * ((size_t)T->nvar + T->nnz) * nc)) return fail(-1002, "out of device memory");
  double *J = h->scratch.as<double>(), *X = J + (size_t)T->nnz * nc;
  CUDA_TRY(cudaMemcpyAsync(J, jvs, sizeof(double) * T->nnz * nc, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(cudaMemcpyAsync(X, x, sizeof(double) * T->nvar * nc, cudaMemcpyHostToDevice, h->stream));
  CUDA_TRY(launch_solve_cells(h->M, ncell, J, X, h->stream));
  CUDA_TRY(cudaMemcpyAsync(x, X, sizeof(double) * T->nvar * nc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int gckpp_gpu_set_keep_active(gckpp_gpu_handle_t *h, int n, const int32_t *idx0)
{
  if (!h || n < 0 || (n > 0 && !idx0)) return fail(-10, "gckpp_gpu_set_keep_active: bad arguments");
  CUDA_TRY(cudaSetDevice(h->device));
  std::vector<unsigned char> m((size_t)h->T->nvar, 0);
  for (int i = 0; i < n; i++) {
    if (idx0[i] < 0 || idx0[i] >= h->T->nvar) return fail(-10, "gckpp_gpu_set_keep_active: index %d out of range", idx0[i]);
    m[idx0[i]] = 1;
  }
  if (h->keep_spc.ensure(m.size())) return fail(-1002, "out of device memory");
  CUDA_TRY(cudaMemcpy(h->keep_spc.p, m.data(), m.size(), cudaMemcpyHostToDevice));
  h->keep_n = n;
  return 0;
}

extern "C" int gckpp_gpu_plan_info(int mech_id, int32_t *info)
{
  const gckpp_host_tables_t *T = host_tables(mech_id);
  const gckpp_sched_tables_t *S = host_sched(mech_id);
  if (!T || !S || !info) return fail(-11, "no shared-memory kernel plan for mechanism %d", mech_id);
  SmemHostPlan p;
  int rc = smem_plan_build(mech_id, T, S, p);
  if (rc) return fail(-11, "plan failed (%d)", rc);
  info[0] = p.s_total; info[1] = (int)(p.stream.size() / 128); info[2] = (int)(p.resident.size() / 128);
  info[3] = (int)p.dir.size(); info[4] = p.n_lu; info[5] = p.n_fwd; info[6] = p.n_bwd; info[7] = SMEM_NC;
  return 0;
}

// Test hook: the per-warp table streams of the warp-group kernel as the host plan lays them out.
//   info[0] = warps per group, info[1] = cells per block, info[2] = shared memory bytes, info[3] = total rows;
//   info[8..8+NPH) = levels per phase; then per warp-stream w (2 + WARP_NSEG + WARP_NPH ints each): w_off, w_rows, seg_off[], nb[].
extern "C" int gckpp_gpu_warp_plan(int mech_id, int32_t *info, int info_cap, uint32_t *stream_out, int64_t stream_cap_words)
{
  const gckpp_host_tables_t *T = host_tables(mech_id);
  const gckpp_wsched_tables_t *S = host_wsched(mech_id);
  if (!T || !S || !info) return fail(-11, "no warp-group kernel plan for mechanism %d", mech_id);
  WarpHostPlan p;
  int rc = warp_plan_build(mech_id, T, S, p);
  if (rc) return fail(-11, "plan failed (%d)", rc);
  const int per = 2 + WARP_NSEG + WARP_NPH;
  if (info_cap < 8 + WARP_NPH + WARP_WG * per) return fail(-10, "info too small");
  int wg = 0;
  for (int w = 0; w < WARP_WG; w++) if (p.w_rows[w] > 0) wg = w + 1;
  info[0] = wg; info[1] = warp_cells_per_block(mech_id); info[2] = warp_smem_bytes(mech_id); info[3] = (int)(p.stream.size() / 128);
  info[4] = WARP_NSEG; info[5] = WARP_NPH; info[6] = WARP_RS; info[7] = per;
  for (int i = 0; i < WARP_NPH; i++) info[8 + i] = p.nlev[i];
  for (int w = 0; w < wg; w++) {
    int32_t *o = info + 8 + WARP_NPH + w * per;
    o[0] = p.w_off[w]; o[1] = p.w_rows[w];
    for (int i = 0; i < WARP_NSEG; i++) o[2 + i] = p.seg_off[w][i];
    for (int i = 0; i < WARP_NPH; i++) o[2 + WARP_NSEG + i] = p.nb[w][i];
  }
  if (stream_out) {
    if ((int64_t)p.stream.size() > stream_cap_words) return fail(-10, "stream buffer too small");
    memcpy(stream_out, p.stream.data(), p.stream.size() * 4);
  }
  return 0;
}

extern "C" int gckpp_gpu_last_stats(gckpp_gpu_handle_t *h, double *stats)
{
  if (!h || !stats) return fail(-10, "gckpp_gpu_last_stats: NULL argument");
  for (int i = 0; i < 16; i++) stats[i] = h->stats[i];
  return 0;
}
