"""Static schedules for the warp-per-cell Rosenbrock kernel (csrc/ros_warp.cu).

One WARP integrates one grid cell; the cell's sparse matrix lives in that warp's slice of shared
memory and nothing is shared between warps, so there are no block barriers: the only ordering
primitive is __syncwarp().  The reference's straight-line / row-sequential sparse code
(KPP/<mech>/gckpp_Function.F90, gckpp_Jacobian.F90, gckpp_LinearAlgebra.F90:46-83 KppDecomp,
:644-2309 KppSolve) is re-expressed as a linear stream of BUNDLES of 32 lane items.  An item is

    target <- f(target_old - sum_k hi_k * lo_k)

in PULL form: every target of a phase appears exactly once, with all of its terms, at the first
dependency level at which all of its operands are final.  A bundle whose operands were written by
an earlier bundle of the same phase carries a SYNC flag (= start of a dependency level).

  vdot  X(i)  = sum coef * A(r)                                    (aggregate form of Fun)
  jvs   G(k)  = G(k) [+ 1/(H*gamma) on the diagonal] - sum coef * B(m)      (G cleared first; the scratch holds
        NREACT doubles, so B(m) is evaluated in two halves: "jvs" uses B(0..NREACT-1), "jvs2" the rest)
  lu    head pivots j < h only (the last m = min(32, NVAR) rows/columns form the TAIL, factorised by
        the kernel in registers):
          L entry (k,c), c < k, c < h :  G = (G - sum_j L(k,j) U(j,c)) * rinv(c)
          U entry (k,c), c > k        :  G =  G - sum_j L(k,j) U(j,c)
          diagonal (k,k), k < h       :  G = 1 / (G - sum ...)        (reciprocal pivot, singular test)
          tail block (k,c >= h)       :  G =  G - sum_{j<h} L(k,j) U(j,c)   (Schur complement)
  fwd   rows with L entries in head columns: X(i) = X(i) - sum_{j<h} L(i,j) X(j)
  bwd   head rows (after the tail chains): X(i) = (X(i) - sum_c U(i,c) X(c)) * rinv(i)

Rows with many terms are split over g = 2^s adjacent lanes whose partial sums are combined with a
segmented shuffle; sums are therefore re-associated with respect to the generated order
(differences at rounding level; ros_generic.cu keeps the reference order).

Table encoding, per lane a sequence of 16-byte rows (4 x uint32):
  row 0 of a bundle = (hdr, meta, t0, t1); further rows = (t2..t5), (t6..t9), ...
      rows = 1 + ceil(max(T - 2, 0) / 4);  the kernel executes exactly T terms
  hdr  = target byte offset | aux byte offset << 16      (aux: the reciprocal pivot an L entry / a bwd row is scaled by;
         G slot NNZ+1 holds 1.0 for targets that are not scaled; lanes that write nothing target the 0.0 slot of the
         array (G slot NNZ, X slot XPAD), so the kernel loads target and scale unconditionally)
  meta = T | lg << 6 | SYNC << 9 | WRITE << 10 | MUL << 11 | DIAG << 12 | UDIAG << 18     (T, lg, SYNC, UDIAG uniform over
         the bundle; bits 13-17 are filled in by the host plan)
  t    = hi << 16 | lo, BYTE offsets into the arrays of the phase:
           vdot/jvs: hi -> coefficient table, lo -> A / B scratch
           lu      : hi -> G (L(k,j)),        lo -> G (U(j,c))
           fwd/bwd : hi -> G,                 lo -> X
      lanes with fewer terms than T are padded with a term whose product is an exact zero.
Lane placement and the order of a lane's terms are chosen to spread the 16 addresses of each
half-warp access over the 16 eight-byte shared-memory banks.
"""
import os
import numpy as np

LMAX = int(os.environ.get("GCKPP_W_LMAX", 8))      # terms per lane before a row is split over more lanes
TAIL = 32
NONE = 0xFFFF

F_SYNC, F_WRITE, F_MUL, F_DIAG = 1 << 9, 1 << 10, 1 << 11, 1 << 12
F_UDIAG = 1 << 18          # uniform over the bundle: some lane has F_DIAG
PHASES = ["vdot", "jvs", "jvs2", "lu", "fwd", "bwd"]


def _pow2ceil(x):
    g = 1
    while g < x:
        g *= 2
    return g


class Bundle:
    __slots__ = ("T", "lg", "sync", "hdr", "flags", "terms", "nreal")
    # hdr[32], flags[32] (lane flags), terms[32][T] (already padded)


def _bank(off):
    return (off >> 3) & 15


def _order_terms(lanes_terms, T, pad):
    """lanes_terms: 32 lists of (hi, lo).  Returns 32 lists of length T (padded), ordered so that at
    every term position the addresses of a half-warp fall into distinct 8-byte banks where possible
    (equal addresses are broadcasts and do not conflict)."""
    out = [[None] * T for _ in range(32)]
    remaining = [list(t) for t in lanes_terms]
    for k in range(T):
        for half in (0, 16):
            load_hi = {}
            load_lo = {}
            # lanes with the fewest choices first
            lanes = sorted(range(half, half + 16), key=lambda l: len(remaining[l]))
            for l in lanes:
                rem = remaining[l]
                if not rem:
                    out[l][k] = pad
                    continue
                best, bc = 0, None
                for idx, (hi, lo) in enumerate(rem):
                    bh, bl = _bank(hi), _bank(lo)
                    sh = load_hi.get(bh, set())
                    sl = load_lo.get(bl, set())
                    c = (0 if hi in sh else len(sh)) + (0 if lo in sl else len(sl))
                    if bc is None or c < bc:
                        best, bc = idx, c
                        if c == 0:
                            break
                hi, lo = rem.pop(best)
                load_hi.setdefault(_bank(hi), set()).add(hi)
                load_lo.setdefault(_bank(lo), set()).add(lo)
                out[l][k] = (hi, lo)
    for l in range(32):
        assert not remaining[l]
    return out


def conflict_degree(addrs):
    """wavefronts of one 8-byte shared-memory access of a warp (two half-warp phases)"""
    tot = 0
    for half in (0, 16):
        banks = {}
        for a in addrs[half:half + 16]:
            banks.setdefault(_bank(a), set()).add(a)
        tot += max(len(s) for s in banks.values())
    return tot


class Packer:
    def __init__(self, pad, tzero, one, lmax=LMAX, optimise=True):
        self.bundles = []
        self.pad = pad
        self.idle_hdr = tzero | (one << 16)     # lanes that write nothing: target = the 0.0 slot, scale = the 1.0 slot
        self.one = one
        self.lmax = lmax
        self.optimise = optimise

    def add_level(self, items):
        """items: list of (target_off, aux_off, lane_flags, [(hi, lo), ...]); one dependency level"""
        its = []
        for tgt, aux, fl, terms in items:
            n = len(terms)
            g = min(32, _pow2ceil((n + self.lmax - 1) // self.lmax)) if n > 0 else 1
            its.append((g, (n + g - 1) // g, tgt, aux, fl, terms))
        its.sort(key=lambda t: (-t[0], -t[1], t[2]))
        first = True
        i = 0
        while i < len(its):
            G = its[i][0]
            slots = 32 // G
            chunk = its[i:i + slots]
            i += slots
            lanes_terms = [[] for _ in range(32)]
            hdr = [self.idle_hdr] * 32
            flags = [0] * 32
            for s, (g, _, tgt, aux, fl, terms) in enumerate(chunk):
                n = len(terms)
                per = (n + G - 1) // G if n else 0
                for p in range(G):
                    lane = s * G + p
                    lanes_terms[lane] = terms[p * per:(p + 1) * per] if per else []
                    if p == 0:
                        hdr[lane] = tgt | ((aux if fl & F_MUL else self.one) << 16)
                        flags[lane] = fl | F_WRITE
            T = max(len(t) for t in lanes_terms)
            assert T < 64
            b = Bundle()
            b.T = T
            b.lg = G.bit_length() - 1
            b.sync = first
            first = False
            b.hdr = hdr
            if any(f & F_DIAG for f in flags):
                flags = [f | F_UDIAG for f in flags]
            b.flags = flags
            b.nreal = sum(len(t) for t in lanes_terms)
            if self.optimise and T > 0:
                b.terms = _order_terms(lanes_terms, T, self.pad)
            else:
                b.terms = [list(t) + [self.pad] * (T - len(t)) for t in lanes_terms]
            self.bundles.append(b)


def bundle_rows(b):
    """-> uint32 [nrows, 32, 4]"""
    nrows = 1 + (max(b.T - 2, 0) + 3) // 4
    out = np.zeros((nrows, 32, 4), np.uint32)
    for l in range(32):
        meta = b.T | (b.lg << 6) | (F_SYNC if b.sync else 0) | b.flags[l]
        seq = [b.hdr[l], meta] + [(hi << 16) | lo for hi, lo in b.terms[l]]
        seq += [0] * (nrows * 4 - len(seq))
        out[:, l, :] = np.array(seq, np.uint32).reshape(nrows, 4)
    return out


def split_bundles(rows):
    """[(first row, number of rows, meta of lane 0)] of the bundles in a row array"""
    out = []
    r = 0
    while r < rows.shape[0]:
        meta = int(rows[r, 0, 1])
        T = meta & 63
        n = 1 + (max(T - 2, 0) + 3) // 4
        out.append((r, n, meta))
        r += n
    return out


def run_rows(rows, hi_arr, lo_arr, tgt_arr, mode, ghinv=0.0):
    """numpy emulation of the kernel's bundle engine on a row array -> (singular flag, bundles executed)"""
    sing = False
    nb = 0
    for r, nrows, _ in split_bundles(rows):
        row0 = rows[r]
        hdr = row0[:, 0].astype(np.int64)
        meta = row0[:, 1].astype(np.int64)
        T = int(meta[0] & 63)
        lg = int((meta[0] >> 6) & 7)
        assert np.all((meta & 63) == T) and np.all(((meta >> 6) & 7) == lg)
        assert np.all((meta & F_SYNC) == (meta[0] & F_SYNC))
        words = rows[r:r + nrows].transpose(1, 0, 2).reshape(32, -1)[:, 2:2 + T].astype(np.int64)
        nb += 1
        a = [np.zeros(32) for _ in range(4)]
        for k in range(T):
            w = words[:, k]
            a[k & 3] = a[k & 3] + hi_arr[(w >> 16) >> 3] * lo_arr[(w & 0xffff) >> 3]
        acc = (a[0] + a[1]) + (a[2] + a[3])
        for s in range(lg):
            sh = np.zeros(32)
            sh[:32 - (1 << s)] = acc[(1 << s):]
            acc = acc + sh
        tg = (hdr & 0xffff) >> 3
        ax = (hdr >> 16) >> 3
        assert np.all(((meta & F_UDIAG) != 0) == bool(np.any(meta & F_DIAG)))
        old = tgt_arr[tg].copy()                 # the kernel loads target and scale of every lane
        for l in range(32):
            if not (meta[l] & F_WRITE):
                assert old[l] == 0.0
                continue
            if mode == "vdot":
                tgt_arr[tg[l]] = acc[l]
            elif mode == "jvs":
                tgt_arr[tg[l]] = (old[l] - acc[l]) + (ghinv if (meta[l] & F_DIAG) else 0.0)
            else:
                gmat = hi_arr                     # lu / solve: hi operands come from G
                assert (meta[l] & F_MUL) or gmat[ax[l]] == 1.0
                v = (old[l] - acc[l]) * gmat[ax[l]]
                if meta[l] & F_DIAG:
                    if not (abs(v) >= np.finfo(np.float64).tiny):
                        sing = True
                    with np.errstate(divide="ignore"):
                        v = 1.0 / v
                tgt_arr[tg[l]] = v
    return sing, nb


class WSchedule:
    def __init__(self, mech, lmax=LMAX, tail=TAIL, optimise=True):
        self.mech = mech
        n = mech.nvar
        crow, diag, icol = mech.lu_crow, mech.lu_diag, mech.lu_icol
        nnz = mech.lu_nonzero
        self.n = n
        self.m = m = min(tail, n)
        self.h = h = n - m
        self.nnz = nnz
        pos = {}
        for i in range(n):
            for p in range(crow[i], crow[i + 1]):
                pos[(i, icol[p])] = p
        Lr = [[c for c in icol[crow[i]:crow[i + 1]] if c < i] for i in range(n)]
        Ur = [[c for c in icol[crow[i]:crow[i + 1]] if c > i] for i in range(n)]
        self.diag = np.array(diag, np.int32)

        # ---- coefficient pool (signed) ----
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
                out.append((coef(c) * 8, i * 8))
            return out

        vd = [(i * 8, 0, 0, coef_terms(mech.Vdot[i], "A")) for i in range(n)]
        dset = set(diag)
        # B(m) is evaluated in two halves through a scratch of NREACT doubles: a target's terms are split by half
        self.nscr = mech.nreact
        assert len(mech.B) <= 2 * self.nscr
        jv, jv2 = [], []
        for k in range(nnz):
            terms = coef_terms(mech.JVS[k], "B")
            t1 = [(c, b) for c, b in terms if b < self.nscr * 8]
            t2 = [(c, b - self.nscr * 8) for c, b in terms if b >= self.nscr * 8]
            if t1 or k in dset:
                jv.append((k * 8, 0, F_DIAG if k in dset else 0, t1))
            if t2:
                jv2.append((k * 8, 0, 0, t2))
        ncoef = len(self.coefs)
        self.coefs.append(0.0)                      # slot ncoef holds 0.0: padding terms of vdot / jvs
        self.coefs = np.array(self.coefs, np.float64)
        PAD_SUM = (ncoef * 8, 0)
        PAD_G = (nnz * 8, nnz * 8)                  # G slot NNZ holds 0.0, slot NNZ+1 holds 1.0
        GZERO, GONE, XZERO = nnz * 8, (nnz + 1) * 8, max(n, 64) * 8
        self.xpad = max(n, 64)                      # X slot XPAD holds 0.0 and is never written (X[0..63] doubles as a buffer)
        PAD_SOLVE = (nnz * 8, self.xpad * 8)
        self.phase = {}

        P = Packer(PAD_SUM, XZERO, GONE, lmax, optimise)
        P.add_level(vd)
        self.phase["vdot"] = P.bundles
        P = Packer(PAD_SUM, GZERO, GONE, lmax, optimise)
        P.add_level(jv)
        self.phase["jvs"] = P.bundles
        P = Packer(PAD_SUM, GZERO, GONE, lmax, optimise)
        if jv2:
            P.add_level(jv2)
        self.phase["jvs2"] = P.bundles

        # ---- LU, head pivots, pull form ----
        lev = {}
        levels = {}
        for k in range(n):
            for p in range(crow[k], crow[k + 1]):
                c = icol[p]
                terms = []
                lv = 0
                for j in Lr[k]:
                    if j >= min(k, c) or j >= h:
                        break
                    q = pos.get((j, c))
                    if q is None:
                        continue
                    terms.append((pos[(k, j)] * 8, q * 8))
                    lv = max(lv, lev.get((k, j), 0), lev.get((j, c), 0))
                fl = 0
                aux = 0
                if c < k and c < h:
                    fl |= F_MUL
                    aux = diag[c] * 8
                    lv = max(lv, lev.get((c, c), 0))
                if c == k and k < h:
                    fl |= F_DIAG
                if not terms and not fl:
                    continue
                lev[(k, c)] = lv + 1
                levels.setdefault(lv + 1, []).append((p * 8, aux, fl, terms))
        P = Packer(PAD_G, GZERO, GONE, lmax, optimise)
        for lv in sorted(levels):
            P.add_level(levels[lv])
        self.phase["lu"] = P.bundles
        self.lu_levels = len(levels)

        # ---- forward sweep: head columns ----
        fl_ = [0] * n
        levels = {}
        for i in range(n):
            terms = [(pos[(i, j)] * 8, j * 8) for j in Lr[i] if j < h]
            if not terms:
                continue
            lv = 1 + max(fl_[j] for j in Lr[i] if j < h)
            if i < h:
                fl_[i] = lv
            levels.setdefault(lv, []).append((i * 8, 0, 0, terms))
        P = Packer(PAD_SOLVE, XZERO, GONE, lmax, optimise)
        for lv in sorted(levels):
            P.add_level(levels[lv])
        self.phase["fwd"] = P.bundles
        self.fwd_levels = len(levels)

        # ---- backward sweep: head rows ----
        bl = [0] * n
        levels = {}
        for i in range(h - 1, -1, -1):
            terms = [(pos[(i, c)] * 8, c * 8) for c in Ur[i]]
            lv = 1 + max([bl[c] for c in Ur[i] if c < h], default=0)
            bl[i] = lv
            levels.setdefault(lv, []).append((i * 8, diag[i] * 8, F_MUL, terms))
        P = Packer(PAD_SOLVE, XZERO, GONE, lmax, optimise)
        for lv in sorted(levels):
            P.add_level(levels[lv])
        self.phase["bwd"] = P.bundles
        self.bwd_levels = len(levels)

        # tail position table, transposed: tposT[j][i] = position of G(h+i, h+j) or NONE
        self.tposT = np.full((32, 32), NONE, np.uint16)
        for i in range(m):
            for j in range(m):
                p = pos.get((h + i, h + j))
                if p is not None:
                    self.tposT[j, i] = p

    # ---- serialisation ------------------------------------------------------------------------
    def phase_rows(self, name):
        bs = self.phase[name]
        if not bs:
            return np.zeros((0, 32, 4), np.uint32)
        return np.concatenate([bundle_rows(b) for b in bs], axis=0)

    def stats(self):
        out = {}
        for name in PHASES:
            bs = self.phase[name]
            rows = sum(bundle_rows(b).shape[0] for b in bs)
            real = sum(b.nreal for b in bs)
            slots = sum(b.T * 32 for b in bs)
            wf = 0
            for b in bs:
                for k in range(b.T):
                    wf += conflict_degree([b.terms[l][k][0] for l in range(32)])
                    wf += conflict_degree([b.terms[l][k][1] for l in range(32)])
            out[name] = dict(bundles=len(bs), real=real, levels=sum(1 for b in bs if b.sync), rows=rows, slots=slots,
                             wavefronts=wf, ideal_wavefronts=4 * sum(b.T for b in bs),
                             T_hist=sorted({t: sum(1 for b in bs if b.T == t) for t in set(b.T for b in bs)}.items()))
        return out

    # ---- numpy emulation of the kernel's bundle engine (CPU tests) -------------------------------
    def run_phase(self, name, hi_arr, lo_arr, tgt_arr, mode, ghinv=0.0):
        """hi_arr/lo_arr/tgt_arr: float64 arrays indexed by byte offset / 8.  mode in vdot, jvs, lu, solve.
        Returns True if a (near-)zero pivot was met (lu)."""
        sing, nb = run_rows(self.phase_rows(name), hi_arr, lo_arr, tgt_arr, mode, ghinv)
        assert nb == len(self.phase[name])
        return sing

    def emulate_fun(self, A):
        X = self.xbuf(np.zeros(self.n))
        self.run_phase("vdot", self.coefs, A, X, "vdot")
        return X[:self.n]

    def gbuf(self, g=None):
        """the kernel's G array: LU_NONZERO entries, then the 0.0 and the 1.0 slot"""
        G = np.zeros(self.nnz + 2)
        if g is not None:
            G[:self.nnz] = g[:self.nnz]
        G[self.nnz + 1] = 1.0
        return G

    def emulate_jac(self, B, ghinv):
        G = self.gbuf()
        B = np.concatenate([B, np.zeros(2 * self.nscr - len(B))])
        self.run_phase("jvs", self.coefs, B[:self.nscr], G, "jvs", ghinv)
        self.run_phase("jvs2", self.coefs, B[self.nscr:], G, "jvs", ghinv)
        return G

    def _tail_dense(self, G):
        m = self.m
        D = np.zeros((m, m))
        for j in range(m):
            for i in range(m):
                p = self.tposT[j, i]
                if p != NONE:
                    D[i, j] = G[p]
        return D

    def emulate_lu(self, G):
        """in place.  Head: L multipliers, reciprocal diagonal, U unscaled.  Tail block: L multipliers,
        reciprocal diagonal, U scaled by the reciprocal diagonal of its row (the kernel's tail_lu)."""
        sing = self.run_phase("lu", G, G, G, "lu")
        m = self.m
        D = self._tail_dense(G)
        for j in range(m - 1):
            l = np.where(np.arange(m) > j, D[:, j] / D[j, j], 0.0)
            D[:, j] = np.where(np.arange(m) > j, l, D[:, j])
            for c in range(j + 1, m):
                D[:, c] = D[:, c] - l * D[j, c]
        dd = np.diag(D).copy()
        for i in range(m):
            D[i, i] = 1.0 / dd[i]
            D[i, i + 1:] *= D[i, i]
        for j in range(m):
            for i in range(m):
                p = self.tposT[j, i]
                if p != NONE:
                    G[p] = D[i, j]
                else:
                    assert D[i, j] == 0.0
        return G, sing

    def xbuf(self, x):
        """right-hand side in the kernel's X array: slot XPAD holds the 0.0 padding terms read"""
        X = np.zeros(self.xpad + 1)
        X[:self.n] = x
        return X

    def emulate_solve(self, G, x):
        m, h, n = self.m, self.h, self.n
        X = self.xbuf(x)
        self.run_phase("fwd", G, X, X, "solve")
        D = self._tail_dense(G)
        t = X[h:n].copy()
        for j in range(m - 1):
            t = t - np.where(np.arange(m) > j, D[:, j], 0.0) * t[j]
        t = t * np.diag(D)
        for j in range(m - 1, 0, -1):
            t = t - np.where(np.arange(m) < j, D[:, j], 0.0) * t[j]
        X[h:n] = t
        self.run_phase("bwd", G, X, X, "solve")
        assert X[self.xpad] == 0.0
        return X[:n]


if __name__ == "__main__":
    import sys
    from . import ir as IR
    mm = IR.load(sys.argv[1] if len(sys.argv) > 1 else "fullchem")
    s = WSchedule(mm)
    for k, v in s.stats().items():
        print(k, v)
    print("coefs", s.coefs.size, "h", s.h, "m", s.m, "levels lu/fwd/bwd", s.lu_levels, s.fwd_levels, s.bwd_levels)
