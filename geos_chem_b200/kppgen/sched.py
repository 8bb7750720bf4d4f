"""Static parallel schedules for the shared-memory-resident Rosenbrock kernel (csrc/ros_smem.cu).

The reference evaluates Fun / Jac_SP / KppDecomp / KppSolve as straight-line or row-sequential
scalar code (KPP/fullchem/gckpp_Function.F90, gckpp_Jacobian.F90, gckpp_LinearAlgebra.F90:46-83,
:644-2309).  On the GPU one thread block integrates a few cells whose sparse matrix lives in shared
memory, so the same arithmetic is re-expressed as ROUNDS of independent "row items"; a barrier
separates rounds.  This module derives those rounds from the mechanism's sparsity pattern:

  vdot   one round : Vdot(i)  = sum coef * A(r)                     (aggregate form of Fun)
  jvs    one round : G(k)     = -sum coef * B(m)  [+ 1/(H*gamma) on the diagonal]
  lu     two rounds per elimination level (right-looking sparse LU, pivots of one level of the
         elimination DAG are independent):  div: G(k,j) /= G(j,j) ; upd: G(k,c) -= sum_j G(k,j)*G(j,c)
  post   G(i,i) <- 1/G(i,i)  (plain loop in the kernel), then  scale: G(i,c) *= G(i,i)^-1 for c > i
  fwd    one round per level of L (push form): X(i) -= sum_j G(i,j) * X(j)   for the columns j of that level
  bwd    X(i) *= 1/G(i,i) (plain loop), then one round per level of U (push form on the scaled U)

Every round is packed into BUNDLES of 32 lane items.  A row with many terms is split over g = 2^s
adjacent lanes whose partial sums are combined with a segmented warp shuffle.  The sums are therefore
re-associated with respect to the reference's generated order (differences at rounding level; the
table-driven kernel in ros_generic.cu keeps the reference order).

Table encoding (all little-endian uint32):
  terms[bundle.term_base + k*32 + lane]  = (hi << 16) | lo      k < len(lane)
        vdot/jvs: hi = coefficient index, lo = A / B index
        lu upd  : hi = position of G(k,j), lo = position of G(j,c)
        lu div  : lo = position of the pivot's diagonal
        scale   : lo = position of the row's diagonal
        fwd/bwd : hi = position of G(i,j),  lo = column j
  lanes[b*32 + lane] = (row << 16) | (len << 8) | flags         flags bit0: lane writes the result,
                                                                  bit1: row is a diagonal position (jvs)
  bundles[b] = (term_base, maxlen | log2(g) << 8)
  rounds[r]  = (first bundle, last bundle + 1)
"""
import numpy as np

LMAX = 8          # target number of terms per lane before a row is split over more lanes

PH_VDOT, PH_JVS, PH_LU, PH_SCALE, PH_FWD, PH_BWD = range(6)
PHASE_NAMES = ["vdot", "jvs", "lu", "scale", "fwd", "bwd"]


def _pow2ceil(x):
    g = 1
    while g < x:
        g *= 2
    return g


class Packer:
    def __init__(self, lmax=LMAX):
        self.terms = []      # flat uint32
        self.lanes = []      # flat uint32, 32 per bundle
        self.bundles = []    # (term_base, maxlen | lg << 8)
        self.rounds = []     # (b0, b1, kind)
        self.lmax = lmax

    def add_round(self, items, kind):
        """items: list of (row, [term words], flags). Returns round index."""
        b0 = len(self.bundles)
        its = []
        for row, tw, fl in items:
            n = len(tw)
            g = min(32, _pow2ceil((n + self.lmax - 1) // self.lmax)) if n > 0 else 1
            its.append((g, -((n + g - 1) // g), row, tw, fl))
        its.sort(key=lambda t: (-t[0], t[1]))
        i = 0
        while i < len(its):
            G = its[i][0]
            slots = 32 // G
            chunk = its[i:i + slots]
            i += slots
            lanes = [0] * 32
            pieces = [[] for _ in range(32)]
            for s, (g, _, row, tw, fl) in enumerate(chunk):
                # split over G lanes (G >= g): contiguous, nearly equal pieces
                n = len(tw)
                per = (n + G - 1) // G if n else 0
                for p in range(G):
                    pc = tw[p * per:(p + 1) * per] if per else []
                    lane = s * G + p
                    pieces[lane] = pc
                    f = (fl | 1) if p == 0 else (fl & ~1)
                    lanes[lane] = (row << 16) | (len(pc) << 8) | (f & 0xff)
                    assert len(pc) < 256 and row < 65536
            maxlen = max(len(p) for p in pieces)
            base = len(self.terms)
            for k in range(maxlen):
                for lane in range(32):
                    self.terms.append(pieces[lane][k] if k < len(pieces[lane]) else 0)
            lg = G.bit_length() - 1
            self.bundles.append((base, maxlen | (lg << 8)))
            self.lanes.extend(lanes)
        self.rounds.append((b0, len(self.bundles), kind))
        return len(self.rounds) - 1


class Schedule:
    """All rounds of one mechanism + the phase directory."""

    def __init__(self, mech, lmax=LMAX):
        self.mech = mech
        n = mech.nvar
        crow, diag, icol = mech.lu_crow, mech.lu_diag, mech.lu_icol
        self.n = n
        pos = {}
        for i in range(n):
            for p in range(crow[i], crow[i + 1]):
                pos[(i, icol[p])] = p
        Lr = [[c for c in icol[crow[i]:crow[i + 1]] if c < i] for i in range(n)]
        Ur = [[c for c in icol[crow[i]:crow[i + 1]] if c > i] for i in range(n)]
        Lcol = [[] for _ in range(n)]
        Ucol = [[] for _ in range(n)]
        for i in range(n):
            for j in Lr[i]:
                Lcol[j].append(i)
            for c in Ur[i]:
                Ucol[c].append(i)
        P = Packer(lmax)
        self.phase = {}

        # ---- coefficient pool (signed) ------------------------------------------------------
        self.coefs = []
        cidx = {}

        def coef(txt):
            v = float(txt)
            if v not in cidx:
                cidx[v] = len(self.coefs)
                self.coefs.append(v)
            return cidx[v]

        def coef_terms(terms, kind):
            out = []
            for t in terms:
                fs = t.factors
                if len(fs) == 1 and fs[0][0] == kind:
                    c, i = "1.0", fs[0][1]
                elif len(fs) == 2 and fs[0][0] == "N" and fs[1][0] == kind:
                    c, i = fs[0][1], fs[1][1]
                else:
                    raise ValueError("unexpected term %r" % (t,))
                if t.neg:
                    c = "-" + c
                out.append((coef(c) << 16) | i)
            return out

        # ---- vdot ---------------------------------------------------------------------------
        r0 = len(P.rounds)
        P.add_round([(i, coef_terms(mech.Vdot[i], "A"), 0) for i in range(n)], PH_VDOT)
        self.phase["vdot"] = (r0, len(P.rounds))
        # ---- jvs ----------------------------------------------------------------------------
        r0 = len(P.rounds)
        dset = set(diag)
        P.add_round([(k, coef_terms(mech.JVS[k], "B"), 2 if k in dset else 0) for k in range(mech.lu_nonzero)], PH_JVS)
        self.phase["jvs"] = (r0, len(P.rounds))
        # ---- lu -----------------------------------------------------------------------------
        # Fine-grained DAG of the row-wise elimination (the LU pattern is NOT structurally symmetric,
        # so pivot-row levels alone are not enough): DIV(k,j) waits for every update of G(k,j) and of
        # the pivot G(j,j); UPD(k,j,c) waits for DIV(k,j) and for the final value of G(j,c).
        tfinal = {}
        divs_at, upds_at = {}, {}
        for k in range(n):
            upd_t = {}
            for j in Lr[k]:
                t = max(upd_t.get(j, 0), tfinal.get((j, j), 0))
                divs_at.setdefault(t, []).append((pos[(k, j)], [diag[j]], 0))
                lp = pos[(k, j)]
                for c in Ur[j]:
                    tu = max(t, tfinal.get((j, c), 0))
                    upds_at.setdefault(tu, {}).setdefault(pos[(k, c)], []).append((lp << 16) | pos[(j, c)])
                    upd_t[c] = max(upd_t.get(c, 0), tu + 1)
            for c, t in upd_t.items():
                tfinal[(k, c)] = t
        r0 = len(P.rounds)
        for t in range(max(list(divs_at) + list(upds_at)) + 1):
            if t in divs_at:
                P.add_round(divs_at[t], PH_LU | 0x10)
            if t in upds_at:
                P.add_round([(tg, tw, 0) for tg, tw in sorted(upds_at[t].items())], PH_LU)
        self.phase["lu"] = (r0, len(P.rounds))
        # ---- scale U rows by the reciprocal diagonal ------------------------------------------
        r0 = len(P.rounds)
        P.add_round([(pos[(i, c)], [diag[i]], 0) for i in range(n) for c in Ur[i]], PH_SCALE)
        self.phase["scale"] = (r0, len(P.rounds))
        # ---- forward sweep, push form -------------------------------------------------------------
        fl = [0] * n
        for i in range(n):
            fl[i] = 1 + max([fl[j] for j in Lr[i]], default=-1)
        r0 = len(P.rounds)
        for lev in range(max(fl) + 1):
            tg = {}
            for j in range(n):
                if fl[j] == lev:
                    for i in Lcol[j]:
                        tg.setdefault(i, []).append((pos[(i, j)] << 16) | j)
            if tg:
                P.add_round([(i, tw, 0) for i, tw in sorted(tg.items())], PH_FWD)
        self.phase["fwd"] = (r0, len(P.rounds))
        # ---- backward sweep, push form ----------------------------------------------------------------
        bl = [0] * n
        for i in range(n - 1, -1, -1):
            bl[i] = 1 + max([bl[c] for c in Ur[i]], default=-1)
        r0 = len(P.rounds)
        for lev in range(max(bl) + 1):
            tg = {}
            for c in range(n):
                if bl[c] == lev:
                    for i in Ucol[c]:
                        tg.setdefault(i, []).append((pos[(i, c)] << 16) | c)
            if tg:
                P.add_round([(i, tw, 0) for i, tw in sorted(tg.items())], PH_BWD)
        self.phase["bwd"] = (r0, len(P.rounds))

        self.terms = np.array(P.terms, np.uint32)
        self.lanes = np.array(P.lanes, np.uint32)
        self.bundles = np.array(P.bundles, np.uint32).reshape(-1, 2)
        self.rounds = np.array([(a, b) for a, b, _ in P.rounds], np.uint32).reshape(-1, 2)
        self.round_kind = [k for _, _, k in P.rounds]
        self.coefs = np.array(self.coefs, np.float64)
        self.diag = np.array(diag, np.int32)

    # ---- numpy emulation of the kernel's bundle engine (used by the CPU tests) ------------------
    def run_round(self, r, op, **kw):
        b0, b1 = self.rounds[r]
        for b in range(b0, b1):
            base, ml = self.bundles[b]
            maxlen, lg = int(ml & 0xff), int(ml >> 8)
            lw = self.lanes[b * 32:(b + 1) * 32]
            row = (lw >> 16).astype(np.int64)
            ln = ((lw >> 8) & 0xff).astype(np.int64)
            fl = (lw & 0xff).astype(np.int64)
            acc = np.zeros(32)
            first = np.zeros(32, np.int64)
            for k in range(maxlen):
                w = self.terms[base + k * 32: base + (k + 1) * 32]
                hi = (w >> 16).astype(np.int64)
                lo = (w & 0xffff).astype(np.int64)
                act = k < ln
                if k == 0:
                    first = lo
                if op in ("vdot", "jvs"):
                    acc += np.where(act, self.coefs[hi] * kw["src"][lo], 0.0)
                elif op == "lu":
                    G = kw["G"]
                    acc += np.where(act, G[hi] * G[lo], 0.0)
                elif op in ("fwd", "bwd"):
                    acc += np.where(act, kw["G"][hi] * kw["X"][lo], 0.0)
            g = 1 << lg
            if g > 1:
                acc = acc.reshape(-1, g).sum(axis=1).repeat(g)
            wr = (fl & 1) == 1
            if op == "vdot":
                kw["out"][row[wr]] = acc[wr]
            elif op == "jvs":
                G = kw["G"]
                G[row[wr]] = -acc[wr] + np.where((fl[wr] & 2) != 0, kw["ghinv"], 0.0)
            elif op == "lu":
                G = kw["G"]
                G[row[wr]] = G[row[wr]] - acc[wr]
            elif op == "div":
                G = kw["G"]
                a = wr & (ln > 0)
                G[row[a]] = G[row[a]] / G[first[a]]
            elif op == "scale":
                G = kw["G"]
                a = wr & (ln > 0)
                G[row[a]] = G[row[a]] * G[first[a]]
            elif op in ("fwd", "bwd"):
                X = kw["X"]
                X[row[wr]] = X[row[wr]] - acc[wr]

    def emulate_fun(self, A):
        out = np.zeros(self.n)
        for r in range(*self.phase["vdot"]):
            self.run_round(r, "vdot", src=A, out=out)
        return out

    def emulate_jac(self, B, ghinv):
        G = np.zeros(self.mech.lu_nonzero)
        for r in range(*self.phase["jvs"]):
            self.run_round(r, "jvs", src=B, G=G, ghinv=ghinv)
        return G

    def emulate_lu(self, G):
        """in place: L unit-lower multipliers, diagonal replaced by its reciprocal, U rows scaled"""
        for r in range(*self.phase["lu"]):
            self.run_round(r, "div" if self.round_kind[r] & 0x10 else "lu", G=G)
        G[self.diag] = 1.0 / G[self.diag]
        for r in range(*self.phase["scale"]):
            self.run_round(r, "scale", G=G)
        return G

    def emulate_solve(self, G, X):
        for r in range(*self.phase["fwd"]):
            self.run_round(r, "fwd", G=G, X=X)
        X *= G[self.diag]
        for r in range(*self.phase["bwd"]):
            self.run_round(r, "bwd", G=G, X=X)
        return X

    def stats(self):
        out = {}
        for name, (r0, r1) in self.phase.items():
            nb = [int(self.rounds[r][1] - self.rounds[r][0]) for r in range(r0, r1)]
            words = 0
            for r in range(r0, r1):
                for b in range(*self.rounds[r]):
                    words += int(self.bundles[b][1] & 0xff) * 32 + 32 + 2
            out[name] = dict(rounds=r1 - r0, bundles=sum(nb), bundles_per_round=nb, table_bytes=words * 4)
        return out


if __name__ == "__main__":
    import sys
    from . import ir as IR
    m = IR.load(sys.argv[1] if len(sys.argv) > 1 else "fullchem")
    s = Schedule(m)
    for k, v in s.stats().items():
        print(k, v)
    print("terms", s.terms.size, "lanes", s.lanes.size, "bundles", len(s.bundles), "rounds", len(s.rounds), "coefs", s.coefs.size)
