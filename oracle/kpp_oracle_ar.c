/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see kpp_oracle.h).
 * Auto-reduce Rosenbrock integrator, the restatement of ros_yIntegrator
 * (KPP/fullchem/gckpp_Integrator.F90:789-1237; ros_cPrepareMatrix :1773-1851, ros_cDecomp :1855-1877,
 * cKppDecomp :2624-2661, FunSplitF/FunSplitN :2518-2575, AutoReduce_1stOrder :1702-1712).
 * This file is #included by kpp_oracle.c (it needs ros_ctx_t, ros_method_t and ros_error_norm).
 * The append variant (ros_yIntegratorA, ICNTRL(13)=1) is not restated: it returns -99, which the
 * reference itself treats as "fall back to the standard solver" (:523).
 * Parity status: UNPINNED by the reference (no auto-reduce fixture exists). */

/* keepSpcActive / keepActive of gckpp_Global (set by fullchem_AutoReduce_SetKeepActive and
 * fullchem_AutoReduce_KeepHalogensActive, fullchem_AutoReduceFuncs.F90:40-140) */
static unsigned char g_keep_spc[3][512];
static int g_keep_active[3];

int kpp_oracle_set_keep_active(int mech_id, int n, const int *idx0)
{
  int i;
  if (mech_id < 0 || mech_id > 2) return -1;
  memset(g_keep_spc[mech_id], 0, sizeof g_keep_spc[mech_id]);
  g_keep_active[mech_id] = n > 0;
  for (i = 0; i < n; i++) {
    if (idx0[i] < 0 || idx0[i] >= 512) return -1;
    g_keep_spc[mech_id][idx0[i]] = 1;
  }
  return 0;
}

static int mech_index(const kpp_mech_t *m)
{
  return m == &kpp_mech_fullchem ? 0 : (m == &kpp_mech_Hg ? 1 : 2);
}

/* cKppDecomp (:2624-2661); arrays are 1-based like the Fortran */
static int ckpp_decomp(int rNVAR, const int *cCROW, const int *cDIAG, const int *cICOL, double *JVS, double *W)
{
  int k, kk, j, jj;
  double a;
  for (k = 1; k <= rNVAR; k++) {
    if (fabs(JVS[cDIAG[k]]) < DBL_MIN) return k;
    for (kk = cCROW[k]; kk <= cCROW[k + 1] - 1; kk++) W[cICOL[kk]] = JVS[kk];
    for (kk = cCROW[k]; kk <= cDIAG[k] - 1; kk++) {
      j = cICOL[kk];
      a = -W[j] / JVS[cDIAG[j]];
      W[j] = -a;
      for (jj = cDIAG[j] + 1; jj <= cCROW[j + 1] - 1; jj++) W[cICOL[jj]] = W[cICOL[jj]] + a * JVS[jj];
    }
    for (kk = cCROW[k]; kk <= cCROW[k + 1] - 1; kk++) JVS[kk] = W[cICOL[kk]];
  }
  return 0;
}

int kpp_oracle_ar_integrate(ros_ctx_t *c, const ros_method_t *ros, double *Y, double Tstart, double Tend,
                            double *Tout, const double *AbsTol, const double *RelTol, int Autonomous,
                            int VectorTol, int Max_no_steps, double Roundoff, double Hmin, double Hmax,
                            double Hstart, double FacMin, double FacMax, double FacRej, double FacSafe,
                            double threshold, int target_spc, double thr_ratio, int append)
{
  const kpp_mech_t *m = c->m;
  const int N = m->nvar, S = ros->S, NZ = m->lu_nonzero;
  const double DeltaMin = 1.0E-5;
  const unsigned char *keepSpc = g_keep_spc[mech_index(m)];
  const int keepActive = g_keep_active[mech_index(m)];
  double *buf, *Ynew, *Fcn0, *Fcn, *Prod, *Loss, *LossY, *dFdT, *Yerr, *K, *Jac0, *Ghimj, *cGhimj, *cW;
  int *ibuf, *SPC_MAP, *iSPC_MAP, *cIROW, *cICOL, *JVS_MAP, *cCROW, *cDIAG, *LU_IROW;
  unsigned char *DO_SLV, *DO_JVS;
  double T, H, Hnew, HC, HG, Fac, Err, AR_thr;
  int Direction, i, j, istage, IERR = 0, RejectLastH = 0, RejectMoreH = 0, reduced = 0;
  int rNVAR = N, cNONZERO = NZ;
  int *IST = c->ISTATUS;
  double *RST = c->RSTATUS;

  if (append || !m->fun_is_split) return -99;
  buf = calloc((size_t)N * (9 + S) + (size_t)NZ * 3 + 8, sizeof(double));
  Ynew = buf; Fcn0 = Ynew + N; Fcn = Fcn0 + N; Prod = Fcn + N; Loss = Prod + N; LossY = Loss + N; dFdT = LossY + N;
  Yerr = dFdT + N; cW = Yerr + N; K = cW + N + 1; Jac0 = K + (size_t)N * S; Ghimj = Jac0 + NZ; cGhimj = Ghimj + NZ;
  ibuf = calloc((size_t)N * 4 + (size_t)NZ * 4 + 16, sizeof(int));
  SPC_MAP = ibuf; iSPC_MAP = SPC_MAP + N + 1; cCROW = iSPC_MAP + N + 1; cDIAG = cCROW + N + 3;
  cIROW = cDIAG + N + 3; cICOL = cIROW + NZ + 1; JVS_MAP = cICOL + NZ + 1; LU_IROW = JVS_MAP + NZ + 1;
  DO_SLV = malloc((size_t)N + NZ);
  DO_JVS = DO_SLV + N;
  memset(DO_SLV, 1, (size_t)N + NZ);                   /* DO_SLV = DO_FUN = DO_JVS = .true. (:858-860) */
  for (i = 0; i < N; i++)
    for (j = m->lu_crow[i]; j < m->lu_crow[i + 1]; j++) LU_IROW[j] = i;

  T = Tstart;
  RST[Nhexit] = 0.0;
  H = fmin(fmax(fabs(Hmin), fabs(Hstart)), fabs(Hmax));
  if (fabs(H) <= 10.0 * Roundoff) H = DeltaMin;
  Direction = (Tend >= Tstart) ? +1 : -1;
  H = Direction * H;
  /* K = 0, Ghimj = 0 (:877-878): calloc */

  while ((Direction > 0 && ((T - Tend) + Roundoff <= 0.0)) ||
         (Direction < 0 && ((Tend - T) + Roundoff <= 0.0))) {
    if (IST[Nstp] > Max_no_steps) { IERR = -6; goto done; }
    if (((T + 0.1 * H) == T) || (H <= Roundoff)) { IERR = -7; goto done; }
    H = fmin(H, fabs(Tend - T));
    if (T == Tstart) {           /* FunSplitF: always calculates P, L (:903-905) */
      m->fun_split(Y, c->FIX, c->RCONST, Fcn0, Prod, Loss, c->A, 0);
      for (i = 0; i < N; i++) LossY[i] = Loss[i] * Y[i];
    } else {                     /* FunSplitN */
      m->fun_split(Y, c->FIX, c->RCONST, Fcn0, c->P, c->D, c->A, 0);
    }
    IST[Nfun]++;

    if (!reduced) {              /* :918-1003, 1-based maps */
      int NRMV = 0, Sx = 1, II = 1, III = 1, idx = 0, i1;
      AR_thr = threshold;
      if (target_spc > 0) {
        AR_thr = thr_ratio * fmax(LossY[target_spc - 1], Prod[target_spc - 1]);
        RST[3] = AR_thr;         /* RSTATUS(NARthr) */
      }
      for (i = 1; i <= N; i++) {
        if (!(keepActive && keepSpc[i - 1]) && fabs(LossY[i - 1]) < AR_thr && fabs(Prod[i - 1]) < AR_thr) {
          NRMV++;
          DO_SLV[i - 1] = 0;
          continue;
        }
        SPC_MAP[Sx] = i;
        iSPC_MAP[i] = Sx;
        Sx++;
      }
      rNVAR = N - NRMV;
      for (i1 = 1; i1 <= NZ; i1++) {
        if (DO_SLV[LU_IROW[i1 - 1]] && DO_SLV[m->lu_icol[i1 - 1]]) {
          idx = 1;
          cIROW[1] = iSPC_MAP[LU_IROW[i1 - 1] + 1];
          cICOL[1] = iSPC_MAP[m->lu_icol[i1 - 1] + 1];
          JVS_MAP[1] = i1;
          break;
        }
        DO_JVS[i1 - 1] = 0;
      }
      for (i1 = i1 + 1; i1 <= NZ; i1++) {
        if (DO_SLV[LU_IROW[i1 - 1]] && DO_SLV[m->lu_icol[i1 - 1]]) {
          idx++;
          cIROW[idx] = iSPC_MAP[LU_IROW[i1 - 1] + 1];
          cICOL[idx] = iSPC_MAP[m->lu_icol[i1 - 1] + 1];
          JVS_MAP[idx] = i1;
          if (cIROW[idx] != cIROW[idx - 1]) { II++; cCROW[II] = idx; }
          if (cIROW[idx] == cICOL[idx]) { III++; cDIAG[III] = idx; }
          continue;
        }
        DO_JVS[i1 - 1] = 0;
      }
      cNONZERO = idx;
      cCROW[1] = 1;
      cDIAG[1] = 1;
      cCROW[rNVAR + 1] = cNONZERO + 1;
      cDIAG[rNVAR + 1] = cDIAG[rNVAR] + 1;
      reduced = 1;
    }

    if (!Autonomous) {           /* ros_FunTimeDerivative: rates do not depend on T with ICNTRL(15) = -1 */
      double Delta = sqrt(Roundoff) * fmax(1.0E-6, fabs(T));
      fun_template(c, Y, dFdT);
      IST[Nfun]++;
      waxpy(N, -1.0, Fcn0, dFdT);
      for (i = 0; i < N; i++) dFdT[i] = (1.0 / Delta) * dFdT[i];
    }
    m->jac_sp(Y, c->FIX, c->RCONST, Jac0, c->B, DO_JVS);      /* JacTemplate reacts to DO_JVS */
    IST[Njac]++;

    for (;;) {                   /* UntilAccepted */
      int Nconsecutive = 0, Singular = 1;
      while (Singular) {         /* ros_cPrepareMatrix (:1773-1851) */
        double ghinv;
        int ising;
        for (i = 1; i <= cNONZERO; i++) cGhimj[i] = -Jac0[JVS_MAP[i] - 1];
        ghinv = 1.0 / (Direction * H * ros->Gamma[0]);
        for (i = 1; i <= rNVAR; i++) cGhimj[cDIAG[i]] = cGhimj[cDIAG[i]] + ghinv;
        ising = ckpp_decomp(rNVAR, cCROW, cDIAG, cICOL, cGhimj, cW);
        IST[Ndec]++;
        if (ising == 0) {
          Singular = 0;
        } else {
          IST[Nsng]++;
          Nconsecutive++;
          if (Nconsecutive <= 5) H = H * 0.5;
          else break;
        }
      }
      if (Singular) { IERR = -8; goto done; }
      for (i = 1; i <= cNONZERO; i++) Ghimj[JVS_MAP[i] - 1] = cGhimj[i];   /* back to the full pattern (:1038-1040) */

      for (istage = 1; istage <= S; istage++) {
        double *Ki = K + (size_t)N * (istage - 1);
        if (istage == 1) {
          memcpy(Fcn, Fcn0, sizeof(double) * N);
        } else if (ros->NewF[istage - 1]) {
          memcpy(Ynew, Y, sizeof(double) * N);
          for (j = 1; j <= istage - 1; j++) {
            double alpha_factor = ros->A[(istage - 1) * (istage - 2) / 2 + j - 1];
            for (i = 1; i <= rNVAR; i++)
              Ynew[SPC_MAP[i] - 1] = Ynew[SPC_MAP[i] - 1] + alpha_factor * K[(size_t)N * (j - 1) + SPC_MAP[i] - 1];
          }
          m->fun_split(Ynew, c->FIX, c->RCONST, Fcn, c->P, c->D, c->A, 0);     /* FunSplitN */
          IST[Nfun]++;
        }
        if (istage > 1) {
          HC = ros->C[(istage - 1) * (istage - 2) / 2 + 1 - 1] / (Direction * H);
          for (i = 1; i <= rNVAR; i++) Ki[SPC_MAP[i] - 1] = Fcn[SPC_MAP[i] - 1] + HC * K[SPC_MAP[i] - 1];
        }
        if (istage == 1)
          for (i = 1; i <= rNVAR; i++) Ki[SPC_MAP[i] - 1] = Fcn[SPC_MAP[i] - 1];
        for (j = 2; j <= istage - 1; j++) {
          HC = ros->C[(istage - 1) * (istage - 2) / 2 + j - 1] / (Direction * H);
          for (i = 1; i <= rNVAR; i++)
            Ki[SPC_MAP[i] - 1] = Ki[SPC_MAP[i] - 1] + HC * K[(size_t)N * (j - 1) + SPC_MAP[i] - 1];
        }
        if (!Autonomous && ros->Gamma[istage - 1] != 0.0) {
          HG = Direction * H * ros->Gamma[istage - 1];
          for (i = 1; i <= rNVAR; i++) Ki[SPC_MAP[i] - 1] = Ki[SPC_MAP[i] - 1] + HG * dFdT[SPC_MAP[i] - 1];
        }
        m->solve(Ghimj, Ki, DO_SLV);       /* ros_Solve -> KppSolve reacts to DO_SLV */
        IST[Nsol]++;
      }
      memcpy(Ynew, Y, sizeof(double) * N);
      for (i = 0; i < N; i++) Yerr[i] = 0.0;
      for (j = 1; j <= S; j++)
        for (i = 1; i <= rNVAR; i++) {
          Ynew[SPC_MAP[i] - 1] = Ynew[SPC_MAP[i] - 1] + ros->M[j - 1] * K[(size_t)N * (j - 1) + SPC_MAP[i] - 1];
          Yerr[SPC_MAP[i] - 1] = Yerr[SPC_MAP[i] - 1] + ros->E[j - 1] * K[(size_t)N * (j - 1) + SPC_MAP[i] - 1];
        }
      Err = ros_error_norm(N, Y, Ynew, Yerr, AbsTol, RelTol, VectorTol);
      Fac = fmin(FacMax, fmax(FacMin, FacSafe / pow(Err, 1.0 / ros->ELO)));
      Hnew = H * Fac;
      IST[Nstp]++;
      if ((Err <= 1.0) || (H <= Hmin)) {
        IST[Nacc]++;
        memcpy(Y, Ynew, sizeof(double) * N);
        T = T + Direction * H;
        Hnew = fmax(Hmin, fmin(Hnew, Hmax));
        if (RejectLastH) Hnew = fmin(Hnew, H);
        RST[Nhexit] = H;
        RST[Nhnew] = Hnew;
        RST[Ntexit] = T;
        RejectLastH = 0;
        RejectMoreH = 0;
        H = Hnew;
        break;
      } else {
        if (RejectMoreH) Hnew = H * FacRej;
        RejectMoreH = RejectLastH;
        RejectLastH = 1;
        H = Hnew;
        if (IST[Nacc] >= 1) IST[Nrej]++;
      }
    }
  }
  /* 1st order calculation for removed species per Shen et al. (2020) Eq. 4 (:1222-1232, :1702-1712) */
  for (i = 0; i < N; i++)
    if (!DO_SLV[i]) {
      double P = Prod[i], k = Loss[i], term;
      if (k <= 1.e-30) continue;
      if (Y[i] <= 1.e-30) continue;
      term = P / k;
      Y[i] = term + (Y[i] - term) * exp(-k * (Tend - Tstart));
    }
  IERR = 1;
done:
  *Tout = T;
  free(buf);
  free(ibuf);
  free(DO_SLV);
  return IERR;
}
