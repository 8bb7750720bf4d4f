"""Stream tables of the lane kernel (csrc/ros_lane.cu, csrc/lane_engine.cuh): one grid cell per lane.

The reference's generated straight-line code (KPP/<mech>/gckpp_Function.F90, gckpp_Jacobian.F90) and its sparse
LU / triangular solves (gckpp_LinearAlgebra.F90:46-83, 644-2309) are re-expressed as STREAMS of batches of four
entries.  The control words of a batch are the same for the 32 cells of a warp; the per-cell operand of an entry
(one double of the cell's workspace column) is prefetched 16 batches ahead through a shared-memory ring, and
the randomly accessed vector (the state under evaluation, the working row of the LU, the right-hand side of a
solve) lives in shared memory, [element][lane].

Index space of that vector ("VEC"): 0..NVAR-1 variable species / row entries / solution, NVAR..NSPEC-1 fixed
species, then the literal pool, then ONE (1.0), ZERO (0.0) and DUMMY (scratch target of padding entries).

Matrix storage ("GA", per cell, in stream order of the solves):
  forward region   for i = 0..N-1:  L'(i,j) for j ascending (UNSCALED multipliers: L(i,j) * pivot(j)), then rinv(i) = 1/pivot(i)
  backward region  for i = N-1..0:  U'(i,c) for c ascending (U(i,c) * rinv(i))
so that  forward:  z(i) = (b(i) - sum_j L'(i,j) z(j)) * rinv(i)        backward:  x(i) = z(i) - sum_c U'(i,c) x(c)
and the elimination of row k needs no pivot look-ups:  W(c) -= W(j) * U'(j,c).
(The reference divides by the pivot in the back sweep; here the reciprocal is applied at the end of the forward
sweep -- a re-association at rounding level, like the FMA contraction.)

Streams (records of RECQ uint4; csrc/lane_engine.cuh):
  rates   (2)  entry = 2 words: rcx | v1 << 16, v2 | v3 << 16        T(4t+e) = RCX(rcx) * VEC(v1) * VEC(v2) * VEC(v3)
  sums    (4)  entry word: src | LAST << 16 | DIAG << 17; 4 coefficients      s += coef * T(src); LAST: emit s
  lu      (2)  entry word: q | c << 13 | VALID << 23 | PF << 24; header: kind | j << 3 | FRESH << 13
  solve   (2)  entry word: q | c << 13 | VALID << 23 | LAST << 24 | RINV << 25 | PF << 26; rows of the LAST entries in rec[1]; FRESH in rec[1].z
PF = the operand comes through the ring; without PF it is read from the workspace when the batch is consumed
(its producer is less than 16 batches upstream).  FRESH = the batch reads VEC slots the batch before it writes,
so its operands are loaded after that batch's stores (otherwise one batch early, to overlap the latencies).

numpy emulations of the consumer (same order of operations, same read-before-write pipelining) are used by
tests/test_lsched.py to check the tables against the oracle on the CPU.
"""
import numpy as np

from . import ir as IR
from .emit_cuda import LitPool, term_ints, coef_terms

DEPTH = 16
BATCH = 4
CHUNK = DEPTH            # records per table chunk

K_NOP, K_LOAD, K_ELIM, K_FINL, K_FIND, K_FINU = range(6)
LU_VALID, LU_PF = 1 << 23, 1 << 24
SV_VALID, SV_LAST, SV_RINV, SV_PF = 1 << 23, 1 << 24, 1 << 25, 1 << 26
SM_LAST, SM_DIAG = 1 << 16, 1 << 17
FRESH = 1 << 13


def _pad_records(rec, recq):
    """pad a [nrec][recq*4] uint32 table with zero records to whole chunks; returns (table, nchunk)"""
    n = rec.shape[0]
    nchunk = max(1, -(-n // CHUNK))
    out = np.zeros((nchunk * CHUNK, recq * 4), np.uint32)
    out[:n] = rec
    return out, nchunk


class LaneSchedule:
    def __init__(self, mech):
        m = mech
        self.mech = m
        self.N, self.nspec, self.nreact = m.nvar, m.nspec, m.nreact
        N = self.N
        self.pool = LitPool()
        # ---- rates: A(r) and B(m) terms; literal rates live behind the rate constants in RCX
        a_terms = [self._term(e) for e in m.A]
        b_terms = [self._term(e) for e in m.B] if m.has_jac else []
        self.nlit = len(self.pool.vals)
        self.lit = np.array([float(v) for v in self.pool.vals], np.float64)
        self.ONE = self.nspec + self.nlit
        self.ZERO = self.ONE + 1
        self.DUMMY = self.ONE + 2
        self.nvec = self.ONE + 3
        assert self.nvec < 1024
        self.nrcx = self.nreact + self.nlit
        self.rates_a, self.nchunk_ra = self._rates(a_terms)
        self.na_pad = self.nchunk_ra * CHUNK * BATCH
        # ---- Fun: aggregate sums over A (the split form P - D*V of fullchem's FunTemplate is the same sum re-associated)
        rows = [coef_terms(e, "A") for e in m.Vdot]
        self.sums_v, self.nchunk_sv = self._sums(rows, [False] * N)
        self.has_jac = m.has_jac
        if not m.has_jac:
            return
        self.rates_b, self.nchunk_rb = self._rates(b_terms)
        self.nb_pad = self.nchunk_rb * CHUNK * BATCH
        # ---- matrix layout
        crow, diag, icol = m.lu_crow, m.lu_diag, m.lu_icol
        self.pos = np.full(m.lu_nonzero, -1, np.int64)          # KPP position -> GA position
        col = []
        self.frun, self.brun, self.dslot = [], [None] * N, []
        p = 0
        for i in range(N):
            s = p
            for k in range(crow[i], diag[i]):
                self.pos[k] = p; col.append(icol[k]); p += 1
            self.frun.append((s, p))
            self.pos[diag[i]] = p; col.append(i); self.dslot.append(p); p += 1
        self.nf = p
        for i in range(N - 1, -1, -1):
            s = p
            for k in range(diag[i] + 1, crow[i + 1]):
                self.pos[k] = p; col.append(icol[k]); p += 1
            self.brun[i] = (s, p)
        self.ng = p
        assert self.ng == m.lu_nonzero and self.ng < 8192
        self.col = np.array(col, np.int64)
        # ---- Jac: G = ghinv*I - J written in GA order
        jrows = [None] * self.ng
        isd = [False] * self.ng
        for k in range(m.lu_nonzero):
            jrows[self.pos[k]] = coef_terms(m.JVS[k], "B") if m.JVS[k] else []
        for i in range(N):
            isd[self.dslot[i]] = True
        self.sums_j, self.nchunk_sj = self._sums(jrows, isd)
        self._lu_stream()
        self.fwd, self.nchunk_fwd = self._solve_stream(True)
        self.bwd, self.nchunk_bwd = self._solve_stream(False)

    # ---- builders ----------------------------------------------------------------------------------------------
    def _term(self, e):
        """(rcx, v1, v2, v3) in RCX / VEC index space; literal indices are resolved later (pool still growing)"""
        if e is None:
            return ("L", self.pool.get("0.0"), [])
        assert len(e) == 1
        r, f0, f1, f2 = term_ints(e[0], self.N, self.pool, "A")
        return ("R", r, [f for f in (f0, f1, f2) if f != -1]) if r >= 0 else ("L", ~r, [f for f in (f0, f1, f2) if f != -1])

    def _vidx(self, f):
        return f if f >= 0 else self.nspec + (-2 - f)

    def _rates(self, terms):
        n = len(terms)
        nb = -(-n // BATCH)
        rec = np.zeros((nb, 8), np.uint32)
        for i in range(nb * BATCH):
            if i < n:
                kind, r, fs = terms[i]
                rcx = r if kind == "R" else self.nreact + r
                v = [self._vidx(f) for f in fs] + [self.ONE] * (3 - len(fs))
            else:
                rcx, v = 0, [self.ONE] * 3
            rec[i // BATCH, 2 * (i % BATCH)] = rcx | (v[0] << 16)
            rec[i // BATCH, 2 * (i % BATCH) + 1] = v[1] | (v[2] << 16)
        return _pad_records(rec, 2)

    def _sums(self, rows, isdiag):
        ent = []
        for r, d in zip(rows, isdiag):
            if not r:
                r = [("0.0", 0)]
            for k, (c, i) in enumerate(r):
                ent.append((i, float(c), k == len(r) - 1, d))
        nb = -(-len(ent) // BATCH)
        rec = np.zeros((nb, 16), np.uint32)
        coef = rec[:, 8:].view(np.float64)
        for n, (i, c, last, d) in enumerate(ent):
            rec[n // BATCH, n % BATCH] = i | (SM_LAST if last else 0) | (SM_DIAG if d else 0)
            coef[n // BATCH, n % BATCH] = c
        return _pad_records(rec, 4)

    def _lu_stream(self):
        N = self.N
        batches = []                      # (kind, j, [(q, c)], rows touched) ; entries padded later

        def add(kind, j, ents):
            for s in range(0, max(len(ents), 1), BATCH):
                batches.append([kind, j, ents[s:s + BATCH]])

        for k in range(N):
            f0, f1 = self.frun[k]
            b0, b1 = self.brun[k]
            own = list(range(f0, f1)) + [self.dslot[k]] + list(range(b0, b1))
            add(K_LOAD, 0, [(q, int(self.col[q])) for q in own])
            for p in range(f0, f1):
                j = int(self.col[p])
                u0, u1 = self.brun[j]
                if u1 > u0:
                    add(K_ELIM, j, [(q, int(self.col[q])) for q in range(u0, u1)])
            if f1 > f0:
                add(K_FINL, 0, [(q, int(self.col[q])) for q in range(f0, f1)])
            add(K_FIND, k, [(self.dslot[k], k)])
            if b1 > b0:
                add(K_FINU, 0, [(q, int(self.col[q])) for q in range(b0, b1)])
        nb = len(batches)
        rec = np.zeros((nb, 8), np.uint32)
        wbatch = {}                        # GA position -> batch that wrote it last (FIN*)
        prev_w = set()                     # VEC slots written by the previous batch
        self.lu_stats = dict(batches=nb, fresh=0, direct=0, entries=0)
        for t, (kind, j, ents) in enumerate(batches):
            reads, writes = set(), set()
            if kind in (K_ELIM, K_FIND):
                reads.add(j)
            for e in range(BATCH):
                if e < len(ents):
                    q, c = ents[e]
                    w = q | (c << 13) | LU_VALID
                    if kind in (K_LOAD, K_ELIM):
                        if wbatch.get(q, -10 ** 9) <= t - DEPTH:
                            w |= LU_PF
                        else:
                            self.lu_stats["direct"] += 1
                    if kind == K_LOAD:
                        writes.add(c)
                    elif kind == K_ELIM:
                        reads.add(c); writes.add(c)
                    elif kind in (K_FINL, K_FINU):
                        reads.add(c)
                    self.lu_stats["entries"] += 1
                else:
                    w = (self.DUMMY << 13)
                rec[t, e] = w
            if kind in (K_FINL, K_FIND, K_FINU):
                for q, _ in ents:
                    wbatch[q] = t
            fresh = bool(reads & prev_w)
            self.lu_stats["fresh"] += fresh
            rec[t, 4] = kind | (j << 3) | (FRESH if fresh else 0)
            prev_w = writes
        self.lu, self.nchunk_lu = _pad_records(rec, 2)

    def _solve_stream(self, forward):
        N = self.N
        ents = []                          # (q, c, last, rinv, row)
        rows = range(N) if forward else range(N - 1, -1, -1)
        for i in rows:
            if forward:
                f0, f1 = self.frun[i]
                for q in range(f0, f1):
                    ents.append((q, int(self.col[q]), False, False, i))
                ents.append((self.dslot[i], self.ZERO, True, True, i))
            else:
                b0, b1 = self.brun[i]
                for q in range(b0, b1):
                    ents.append((q, int(self.col[q]), q == b1 - 1, False, i))
        # batches: an entry may not read an x that a LAST entry of the same batch writes
        batches, cur, wr = [], [], set()
        for en in ents:
            if len(cur) == BATCH or en[1] in wr or (en[2] and en[4] in wr):
                batches.append(cur); cur, wr = [], set()
            cur.append(en)
            if en[2]:
                wr.add(en[4])
        if cur:
            batches.append(cur)
        nb = len(batches)
        rec = np.zeros((nb, 8), np.uint32)
        prev_w = set()
        st = dict(batches=nb, fresh=0, entries=len(ents))
        for t, b in enumerate(batches):
            reads, writes = set(), set()
            rws = [0, 0, 0, 0]
            for e in range(BATCH):
                if e < len(b):
                    q, c, last, rinv, i = b[e]
                    rec[t, e] = q | (c << 13) | SV_VALID | SV_PF | (SV_LAST if last else 0) | (SV_RINV if rinv else 0)
                    reads.add(c)
                    if last:
                        reads.add(i); writes.add(i); rws[e] = i
                else:
                    rec[t, e] = self.ZERO << 13
            rec[t, 4] = rws[0] | (rws[1] << 16)
            rec[t, 5] = rws[2] | (rws[3] << 16)
            fresh = bool(reads & prev_w)
            st["fresh"] += fresh
            rec[t, 6] = FRESH if fresh else 0
            prev_w = writes
        if forward:
            self.fwd_stats = st
        else:
            self.bwd_stats = st
        return _pad_records(rec, 2)

    # ---- numpy emulation of the consumers ------------------------------------------------------------------------
    def vec(self, y):
        """VEC with the state y[0:nspec] loaded"""
        v = np.zeros(self.nvec)
        v[:self.nspec] = y
        v[self.nspec:self.nspec + self.nlit] = self.lit
        v[self.ONE] = 1.0
        return v

    def rcx(self, rconst):
        return np.concatenate([np.asarray(rconst, np.float64), self.lit])

    def emulate_rates(self, tab, vec, rcx, nout):
        out = np.zeros(tab.shape[0] * BATCH)
        for t in range(tab.shape[0]):
            for e in range(BATCH):
                w0, w1 = int(tab[t, 2 * e]), int(tab[t, 2 * e + 1])
                out[t * BATCH + e] = rcx[w0 & 0xffff] * vec[w0 >> 16] * vec[w1 & 0xffff] * vec[w1 >> 16]
        return out

    def emulate_sums(self, tab, src, ghinv=0.0, negate=False):
        out = []
        coef = np.ascontiguousarray(tab[:, 8:]).view(np.float64)
        s = 0.0
        for t in range(tab.shape[0]):
            for e in range(BATCH):
                w = int(tab[t, e])
                s = s + coef[t, e] * src[w & 0xffff]
                if w & SM_LAST:
                    out.append(((ghinv if w & SM_DIAG else 0.0) - s) if negate else s)
                    s = 0.0
        return np.array(out)

    def emulate_fun(self, y, rconst):
        A = self.emulate_rates(self.rates_a, self.vec(y), self.rcx(rconst), self.nreact)
        return self.emulate_sums(self.sums_v, A)

    def emulate_jac(self, y, rconst, ghinv):
        B = self.emulate_rates(self.rates_b, self.vec(y), self.rcx(rconst), len(self.mech.B))
        return self.emulate_sums(self.sums_j, B, ghinv, True)

    def emulate_lu(self, ga):
        """in place on GA; returns singular flag.  Follows the kernel's pipelining: the ring operand of a batch is the
        value GA held when the batch 16 upstream had been consumed; the VEC operands of a batch that is not FRESH
        are read before the previous batch stores."""
        tab = self.lu
        nb = tab.shape[0]
        W = np.zeros(self.nvec)
        ring = {}
        sing = False
        rinv = 0.0

        def issue(t):
            if t < nb:
                ring[t] = [ga[int(tab[t, e]) & 0x1fff] if int(tab[t, e]) & LU_PF else None for e in range(BATCH)]

        def preload(t):
            h = int(tab[t, 4])
            kind, j = h & 7, (h >> 3) & 1023
            ops = []
            for e in range(BATCH):
                w = int(tab[t, e])
                q, c = w & 0x1fff, (w >> 13) & 1023
                g = ring[t][e]
                if g is None and (w & LU_VALID) and kind in (K_LOAD, K_ELIM):
                    g = ga[q]
                ops.append((w, q, c, g, W[c]))
            return kind, j, W[j], ops

        for t in range(min(DEPTH, nb)):
            issue(t)
        pre = preload(0)
        for t in range(nb):
            nxt = None
            if t + 1 < nb and not (int(tab[t + 1, 4]) & FRESH):
                nxt = preload(t + 1)
            kind, j, wj, ops = pre
            for w, q, c, g, wc in ops:
                if kind == K_LOAD:
                    if w & LU_VALID:
                        W[c] = g
                elif kind == K_ELIM:
                    if w & LU_VALID:
                        W[c] = wc - wj * g
                elif kind == K_FINL:
                    if w & LU_VALID:
                        ga[q] = wc
                elif kind == K_FINU:
                    if w & LU_VALID:
                        ga[q] = wc * rinv
            if kind == K_FIND:
                sing |= not (abs(wj) >= np.finfo(np.float64).tiny)
                with np.errstate(divide="ignore"):
                    rinv = 1.0 / wj
                ga[ops[0][1]] = rinv
            issue(t + DEPTH)
            if t + 1 < nb and nxt is None:
                nxt = preload(t + 1)
            pre = nxt
            ring.pop(t, None)
        return sing

    def emulate_solve(self, ga, b):
        x = np.zeros(self.nvec)
        x[:self.N] = b
        for tab in (self.fwd, self.bwd):
            nb = tab.shape[0]

            def preload(t):
                ops = []
                rws = (int(tab[t, 4]) & 0xffff, int(tab[t, 4]) >> 16, int(tab[t, 5]) & 0xffff, int(tab[t, 5]) >> 16)
                for e in range(BATCH):
                    w = int(tab[t, e])
                    ops.append((w, ga[w & 0x1fff] if w & SV_PF else 0.0, x[(w >> 13) & 1023], rws[e], x[rws[e]]))
                return ops

            s = 0.0
            pre = preload(0)
            for t in range(nb):
                nxt = None
                if t + 1 < nb and not (int(tab[t + 1, 6]) & FRESH):
                    nxt = preload(t + 1)
                for w, g, xc, i, xi in pre:
                    if not (w & SV_VALID):
                        continue
                    if not (w & SV_RINV):
                        s = s + g * xc
                    if w & SV_LAST:
                        v = xi - s
                        if w & SV_RINV:
                            v = v * g
                        x[i] = v
                        s = 0.0
                if t + 1 < nb and nxt is None:
                    nxt = preload(t + 1)
                pre = nxt
        return x[:self.N].copy()

    def ga_from_kpp(self, jvs):
        ga = np.zeros(self.ng)
        ga[self.pos] = jvs
        return ga


def build(name):
    return LaneSchedule(IR.load(name))
