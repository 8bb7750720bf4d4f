"""Static parallel schedules for the shared-memory-resident Rosenbrock kernel (csrc/ros_smem.cu).

The reference evaluates Fun / Jac_SP / KppDecomp / KppSolve as straight-line or row-sequential
scalar code (KPP/fullchem/gckpp_Function.F90, gckpp_Jacobian.F90, gckpp_LinearAlgebra.F90:46-83,
:644-2309).  On the GPU one thread block integrates a few cells whose sparse matrix lives in shared
memory, so the same arithmetic is re-expressed as ROUNDS of independent "row items"; a barrier
separates rounds.  The matrix is split into a HEAD (rows/columns < h) and a dense-ish TAIL (the last
m = min(32, NVAR) rows/columns, where KPP's ordering concentrates the fill-in): the elimination DAG of
the head is shallow and wide (rounds), the tail is a chain (one warp per cell, registers + shuffles).

  vdot   one round : Vdot(i)  = sum coef * A(r)                     (aggregate form of Fun)
  jvs    one round : G(k)     = -sum coef * B(m)  [+ 1/(H*gamma) on the diagonal]
  lu     head pivots j < h, right-looking, levels of the fine-grained task DAG:
             div: G(k,j) /= G(j,j)        upd: G(k,c) -= sum_j G(k,j)*G(j,c)     (all k > j, c > j)
         then the m x m Schur complement is factorised by the tail code
  post   row i (one thread): G(i,i) <- 1/G(i,i), then G(i,c) *= G(i,i) for c > i
  fwd    push form, one round per level of the head of L: X(i) -= sum_j G(i,j) * X(j), head columns j
         of that level, all rows i (head and tail); then the tail chain
  bwd    X(i) *= 1/G(i,i); tail chain; round 0 pushes the tail columns into the head rows, then one
         round per level of the head of (scaled) U

Every round is packed into BUNDLES of 32 lane items.  A row with many terms is split over g = 2^s
adjacent lanes whose partial sums are combined with a segmented warp shuffle.  The sums are therefore
re-associated with respect to the reference's generated order (differences at rounding level; the
table-driven kernel in ros_generic.cu keeps the reference order).

Table encoding: per lane a sequence of 16-byte CHUNKS (4 x uint32):
  chunk 0 of a bundle = (lw, t0, t1, t2), further chunks = (t3..t6), ...   nchunks = 1 + ceil((maxlen-3)/4)
  lw  = row | len << 13 | maxlen << 19 | log2(g) << 25 | flags << 28
        flags bit0: lane writes the result, bit1: row is a diagonal position (jvs)
  t   = hi << 16 | lo, both BYTE offsets (8 * index) so the kernel adds them to an array base directly
        vdot/jvs: hi = coefficient, lo = A / B entry
        lu upd  : hi = G(k,j), lo = G(j,c)
        lu div  : lo = the pivot's diagonal
        fwd/bwd : hi = G(i,j),  lo = X(j)
      Lanes with fewer terms than the bundle's maxlen are padded with a term whose product is an exact
      zero (coefficient slot NCOEF holds 0.0; G slot LU_NONZERO holds 0.0), so the kernel needs no
      per-term predicate.
A round uses W = min(NW, bundles) warps; bundle b of the round goes to warp b % W.
"""
import numpy as np

import os as _os
LMAX = int(_os.environ.get('GCKPP_LMAX', 8))          # target number of terms per lane before a row is split over more lanes
SOLVE_LMAX = int(_os.environ.get('GCKPP_SOLVE_LMAX', 7))    # triangular sweeps: at most two chunks per bundle (3 + 4 terms), both prefetched before the barrier
EXACT_STEPS = int(_os.environ.get('GCKPP_EXACT_STEPS', 0))   # bit mask of the ops whose bundles apply exactly maxlen steps: 1 vdot, 2 jvs, 4 lu, 8 sweeps  # the kernel applies maxlen steps per bundle (SMEM_EXACT_STEPS), not whole chunks
RECIP_DIAG = int(_os.environ.get('GCKPP_RECIP_DIAG', 1))    # head pivots: reciprocal stored when the diagonal is final, L entries multiplied by it
BANK_GROUP = int(_os.environ.get('GCKPP_BANK_GROUP', 1))    # ... and choose which items share a half-warp
BANK_OPT = int(_os.environ.get('GCKPP_BANK_OPT', 1))      # place the terms of a bundle against shared-memory bank conflicts
TAIL = 32         # tail block size (one lane per tail row)
NONE = 0xFFFF

PHASE_NAMES = ["vdot", "jvs", "lu", "scale", "fwd", "bwd"]
K_DIV = 0x10


def _pow2ceil(x):
    g = 1
    while g < x:
        g *= 2
    return g


class Bundle:
    __slots__ = ("lw", "pieces", "maxlen", "pad")


class Packer:
    def __init__(self, lmax=LMAX):
        self.bundles = []    # Bundle
        self.rounds = []     # (b0, b1, kind)
        self.lmax = lmax

    def add_round(self, items, kind, lmax=None, pad=0):
        """items: list of (row, [term words], flags); pad = the no-op term word of this round's operation"""
        lmax = lmax or self.lmax
        b0 = len(self.bundles)
        its = []
        for row, tw, fl in items:
            n = len(tw)
            g = min(32, _pow2ceil((n + lmax - 1) // lmax)) if n > 0 else 1
            its.append((g, -((n + g - 1) // g), row, tw, fl))
        its.sort(key=lambda t: (-t[0], t[1]))
        if BANK_OPT and BANK_GROUP and not (kind & K_DIV):
            its = _group_by_banks(its)
        i = 0
        while i < len(its):
            G = its[i][0]
            slots = 32 // G
            chunk = its[i:i + slots]
            i += slots
            pieces = [[] for _ in range(32)]
            meta = [(0, 0)] * 32
            for s, (g, _, row, tw, fl) in enumerate(chunk):
                n = len(tw)
                per = (n + G - 1) // G if n else 0
                for p in range(G):
                    pc = tw[p * per:(p + 1) * per] if per else []
                    lane = s * G + p
                    pieces[lane] = pc
                    f = (fl | 1) if p == 0 else (fl & ~1)
                    meta[lane] = (row, f & 7)
            maxlen = max(len(p) for p in pieces)
            assert maxlen < 64
            lg = G.bit_length() - 1
            b = Bundle()
            b.maxlen = maxlen
            b.pad = pad
            b.pieces = pieces
            b.lw = [meta[l][0] | (len(pieces[l]) << 13) | (maxlen << 19) | (lg << 25) | (meta[l][1] << 28) for l in range(32)]
            assert all(meta[l][0] < 8192 for l in range(32))
            if BANK_OPT and not (kind & K_DIV):
                bank_optimize(b, kind)
            self.bundles.append(b)
        self.rounds.append((b0, len(self.bundles), kind))


# ---- shared-memory bank model -------------------------------------------------------------------------------------
# The bundle engine gathers two 8-byte operands per term and cell with one LDS.64 per (term step, cell, operand): all
# 32 lanes of the instruction use the same array base, so only the lanes' own offsets decide the conflicts.  A 64-bit
# shared load is served half-warp by half-warp; within a half-warp two lanes collide when their words fall into the
# same pair of banks ((offset / 8) mod 16) at different addresses, and the instruction takes max-multiplicity
# wavefronts per half-warp.  ncu of the round-1 kernel: 40 % of all shared-memory wavefronts were such replays and the
# LSU data pipe was ~50 % busy, so the order in which a lane applies its terms -- free, the sums are re-associated
# anyway -- is chosen to keep the lanes of a half-warp on different bank pairs (bank_optimize, below).
def _bp(off):
    return (off >> 3) & 15


def _exact(kind):
    return bool((EXACT_STEPS >> {0: 0, 1: 1, 2: 2}.get(kind & 7, 3)) & 1)


def bundle_wavefronts(b, kind=4):
    """model: wavefronts of the operand gathers of one bundle and one cell (2 per instruction = conflict-free)"""
    nch = 1 + max(0, (b.maxlen - 3 + 3) // 4)
    nstep = b.maxlen if _exact(kind) else nch * 4 - 1
    tot = 0
    for s in range(nstep):
        for half in (0, 16):
            for sh in (16, 0):
                occ = {}
                for l in range(half, half + 16):
                    w = b.pieces[l][s] if s < len(b.pieces[l]) else b.pad
                    off = (w >> sh) & 0xffff
                    occ.setdefault(_bp(off), set()).add(off)
                tot += max(len(v) for v in occ.values())
    return tot, nstep * 4


def bank_optimize(b, kind=4, passes=3):
    """Re-place the terms of every lane over the bundle's term steps (positions beyond a lane's own terms hold the no-op
    pad word) so that, step by step, the lanes of a half-warp hit different bank pairs.  Greedy over the lanes, longest
    first, each lane an assignment problem (terms x steps) against the lanes already placed; a few refinement passes."""
    from scipy.optimize import linear_sum_assignment
    nch = 1 + max(0, (b.maxlen - 3 + 3) // 4)
    nstep = b.maxlen if _exact(kind) else nch * 4 - 1
    if b.maxlen == 0:
        return
    for half in (0, 16):
        lanes = sorted(range(half, half + 16), key=lambda l: -len(b.pieces[l]))
        place = {l: None for l in lanes}          # lane -> list of nstep words
        # occupancy[s][operand][bank pair] -> {offset: count}
        occ = [[{}, {}] for _ in range(nstep)]

        def add(l, sgn):
            for s, w in enumerate(place[l]):
                for o, sh in enumerate((16, 0)):
                    off = (w >> sh) & 0xffff
                    d = occ[s][o].setdefault(_bp(off), {})
                    d[off] = d.get(off, 0) + sgn
                    if d[off] == 0:
                        del d[off]

        def cost(w, s):
            c = 0
            for o, sh in enumerate((16, 0)):
                off = (w >> sh) & 0xffff
                d = occ[s][o].get(_bp(off))
                if d:
                    c += len(d) - (1 if off in d else 0)         # distinct other addresses on this bank pair
            return c

        def assign(l):
            terms = real[l]
            out = [b.pad] * nstep
            if terms:
                # a step left to the pad word collides like any other address: costs are relative to it
                padc = np.array([cost(b.pad, s) for s in range(nstep)], np.float64)
                C = np.array([[cost(w, s) for s in range(nstep)] for w in terms], np.float64) - padc[None, :]
                ri, ci = linear_sum_assignment(C)
                for r, c_ in zip(ri, ci):
                    out[c_] = terms[r]
            place[l] = out

        real = {l: list(b.pieces[l]) for l in lanes}
        for l in lanes:
            assign(l)
            add(l, +1)
        for _ in range(passes):
            for l in lanes:
                add(l, -1)
                assign(l)
                add(l, +1)
        for l in lanes:
            b.pieces[l] = place[l]
    # every lane now carries nstep words (pads included); maxlen is unchanged, so is the chunk count


def _group_by_banks(its, window=160):
    """Order the single-lane items of a round so that the 16 items of every half-warp spread over the bank pairs: the
    longest remaining item seeds a half-warp, the other 15 come from the next `window` items of the length-sorted list
    (so a bundle's maxlen barely grows), each chosen for the fewest new (bank pair, address) collisions with the
    half-warp so far.  Most LU / sweep bundles carry one or two terms per lane -- no freedom inside the lane, all of it
    in who shares a half-warp.  Items split over several lanes keep their place in front."""
    head = [t for t in its if t[0] > 1]
    pool = [t for t in its if t[0] == 1]
    # lanes left in the bundle the split items end with are filled by the first single-lane items, as before
    out = list(head)
    nhead_lanes = sum(t[0] for t in head) % 32
    if nhead_lanes:
        k = min(len(pool), 32 - nhead_lanes)
        out += pool[:k]
        pool = pool[k:]
    while pool:
        seed = pool.pop(0)
        grp = [seed]
        occ = [{}, {}]

        def put(t):
            for w in t[3]:
                for o, sh in enumerate((16, 0)):
                    off = (w >> sh) & 0xffff
                    occ[o].setdefault(_bp(off), set()).add(off)

        def cost(t):
            c = 0
            for w in t[3]:
                for o, sh in enumerate((16, 0)):
                    off = (w >> sh) & 0xffff
                    d = occ[o].get(_bp(off))
                    if d and off not in d:
                        c += len(d)
            return c

        put(seed)
        while len(grp) < 16 and pool:
            # candidates: the run of items as long as the next one (every extra term step of a bundle is paid by all
            # its lanes, so lengths are never mixed beyond what the plain sorted order would do)
            nl = len(pool[0][3])
            k = 1
            while k < len(pool) and k < window and len(pool[k][3]) == nl:
                k += 1
            win = pool[:k]
            j = min(range(len(win)), key=lambda q: (cost(win[q]), q))
            t = pool.pop(j)
            grp.append(t)
            put(t)
        out += grp
    return out


def bundle_chunks(b):
    """-> uint32 array [nchunks, 32, 4]"""
    nch = 1 + max(0, (b.maxlen - 3 + 3) // 4)
    out = np.zeros((nch, 32, 4), np.uint32)
    for l in range(32):
        seq = [b.lw[l]] + list(b.pieces[l])
        seq += [b.pad] * (nch * 4 - len(seq))
        out[:, l, :] = np.array(seq, np.uint32).reshape(nch, 4)
    return out


class Schedule:
    """All rounds of one mechanism + the phase directory + the tail tables."""

    def __init__(self, mech, lmax=LMAX, tail=TAIL):
        self.mech = mech
        n = mech.nvar
        crow, diag, icol = mech.lu_crow, mech.lu_diag, mech.lu_icol
        self.n = n
        self.m = m = min(tail, n)
        self.h = h = n - m
        pos = {}
        for i in range(n):
            for p in range(crow[i], crow[i + 1]):
                pos[(i, icol[p])] = p
        self.pos = pos
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
                out.append((coef(c) * 8 << 16) | (i * 8))
            return out

        # ---- vdot ---------------------------------------------------------------------------
        nnz = mech.lu_nonzero
        r0 = len(P.rounds)
        vd = [(i, coef_terms(mech.Vdot[i], "A"), 0) for i in range(n)]
        jv = [(k, coef_terms(mech.JVS[k], "B"), 2 if k in set(diag) else 0) for k in range(nnz)
              if mech.JVS[k] or k in set(diag)]
        ncoef = len(self.coefs)               # slot ncoef of the coefficient table holds 0.0
        self.coefs.append(0.0)
        PAD_SUM, PAD_LU, PAD_SOLVE = (ncoef * 8) << 16, ((nnz * 8) << 16) | (nnz * 8), (nnz * 8) << 16
        P.add_round(vd, 0, pad=PAD_SUM)
        self.phase["vdot"] = (r0, len(P.rounds))
        # ---- lu, head pivots ------------------------------------------------------------------
        # Fine-grained DAG of the row-wise elimination (the LU pattern is NOT structurally symmetric,
        # so pivot-row levels alone are not enough): DIV(k,j) waits for every update of G(k,j) and of
        # the pivot G(j,j); UPD(k,j,c) waits for DIV(k,j) and for the final value of G(j,c).
        tfinal = {}
        divs_at, upds_at = {}, {}
        for k in range(n):
            upd_t = {}
            for j in Lr[k]:
                if j >= h:
                    continue
                t = max(upd_t.get(j, 0), tfinal.get((j, j), 0))
                divs_at.setdefault(t, []).append((pos[(k, j)], [diag[j] * 8], 0))
                lp = pos[(k, j)]
                for c in Ur[j]:
                    tu = max(t, tfinal.get((j, c), 0))
                    upds_at.setdefault(tu, {}).setdefault(pos[(k, c)], []).append((lp * 8 << 16) | (pos[(j, c)] * 8))
                    upd_t[c] = max(upd_t.get(c, 0), tu + 1)
            for c, t in upd_t.items():
                tfinal[(k, c)] = t
        # RECIP_DIAG: the diagonal of a head pivot is replaced by its reciprocal by the item that gives it its final value
        # (flag 4: the last LU update of G(j,j), or the Jacobian item when no update touches it), and the "div" rounds
        # multiply by it -- one division per pivot instead of one per L entry.
        inv_at = {}                     # level of the last update -> {diagonal positions}
        inv_jvs = set()
        if RECIP_DIAG:
            for j in range(h):
                t = tfinal.get((j, j), 0)
                if t > 0:
                    inv_at.setdefault(t - 1, set()).add(diag[j])
                else:
                    inv_jvs.add(diag[j])
        self.recip_diag = bool(RECIP_DIAG)
        self.inv_jvs = np.array(sorted(inv_jvs), np.int64)     # diagonals the Jacobian round leaves as reciprocals
        # ---- jvs ----------------------------------------------------------------------------
        r0 = len(P.rounds)
        # structural zeros (LU fill-in slots) are not listed: the kernel clears G before this round
        jv = [(k, tw, fl | (4 if k in inv_jvs else 0)) for k, tw, fl in jv]
        P.add_round(jv, 1, pad=PAD_SUM)
        self.phase["jvs"] = (r0, len(P.rounds))
        r0 = len(P.rounds)
        for t in range(max(list(divs_at) + list(upds_at) + [-1]) + 1):
            if t in divs_at:
                P.add_round(divs_at[t], 2 | K_DIV)
            if t in upds_at:
                assert inv_at.get(t, set()) <= set(upds_at[t])
                P.add_round([(tg, tw, 4 if tg in inv_at.get(t, ()) else 0) for tg, tw in sorted(upds_at[t].items())], 2, pad=PAD_LU)
        assert all(t in upds_at for t in inv_at)
        self.phase["lu"] = (r0, len(P.rounds))
        # ---- scale U rows by the reciprocal diagonal ------------------------------------------
        # (done row-wise by the thread that inverts the diagonal: no table, see the kernel's post-LU pass)
        r0 = len(P.rounds)
        self.phase["scale"] = (r0, len(P.rounds))
        self.crow = np.array(crow, np.int32)
        # ---- forward sweep: head columns, push form -----------------------------------------------
        fl = [0] * n
        for i in range(h):
            fl[i] = 1 + max([fl[j] for j in Lr[i]], default=-1)
        r0 = len(P.rounds)
        for lev in range(max(fl[:h], default=-1) + 1):
            tg = {}
            for j in range(h):
                if fl[j] == lev:
                    for i in Lcol[j]:
                        tg.setdefault(i, []).append((pos[(i, j)] * 8 << 16) | (j * 8))
            if tg:
                P.add_round([(i, tw, 0) for i, tw in sorted(tg.items())], 4, lmax=SOLVE_LMAX, pad=PAD_SOLVE)
        self.phase["fwd"] = (r0, len(P.rounds))
        # ---- backward sweep: tail columns first, then head columns by level ------------------------
        bl = [0] * n
        for i in range(h - 1, -1, -1):
            bl[i] = 1 + max([bl[c] if c < h else -1 for c in Ur[i]], default=-1)
        r0 = len(P.rounds)
        tg = {}
        for c in range(h, n):
            for i in Ucol[c]:
                if i < h:
                    tg.setdefault(i, []).append((pos[(i, c)] * 8 << 16) | (c * 8))
        if tg:
            P.add_round([(i, tw, 0) for i, tw in sorted(tg.items())], 5, lmax=SOLVE_LMAX, pad=PAD_SOLVE)
        for lev in range(max(bl[:h], default=-1) + 1):
            tg = {}
            for c in range(h):
                if bl[c] == lev:
                    for i in Ucol[c]:
                        tg.setdefault(i, []).append((pos[(i, c)] * 8 << 16) | (c * 8))
            if tg:
                P.add_round([(i, tw, 0) for i, tw in sorted(tg.items())], 5, lmax=SOLVE_LMAX, pad=PAD_SOLVE)
        self.phase["bwd"] = (r0, len(P.rounds))

        self.bundles = P.bundles
        self.rounds = P.rounds
        self.coefs = np.array(self.coefs, np.float64)
        self.diag = np.array(diag, np.int32)
        # tail position table, transposed: tposT[j][i] = position of G(h+i, h+j) or NONE
        self.tposT = np.full((m, m), NONE, np.uint16)
        for i in range(m):
            for j in range(m):
                p = pos.get((h + i, h + j))
                if p is not None:
                    self.tposT[j, i] = p

    # ---- serialisation for the kernel ------------------------------------------------------------
    def chunk_table(self):
        """(chunks uint32 [nrows,32,4], bundle_row uint32 [nb+1])"""
        rows = []
        off = [0]
        for b in self.bundles:
            c = bundle_chunks(b)
            rows.append(c)
            off.append(off[-1] + c.shape[0])
        return np.concatenate(rows, axis=0), np.array(off, np.uint32)

    # ---- numpy emulation of the kernel's bundle engine (used by the CPU tests) ------------------
    def run_round(self, r, op, **kw):
        b0, b1, _ = self.rounds[r]
        for b in range(b0, b1):
            ch = bundle_chunks(self.bundles[b])                   # exercise the packed form
            words = ch.transpose(1, 0, 2).reshape(32, -1)          # [lane][lw, t0, t1, ...]
            lw = words[:, 0]
            row = (lw & 0x1fff).astype(np.int64)
            ln = ((lw >> 13) & 0x3f).astype(np.int64)
            maxlen = int((lw[0] >> 19) & 0x3f)
            lg = int((lw[0] >> 25) & 7)
            fl = ((lw >> 28) & 7).astype(np.int64)
            assert ch.shape[0] == 1 + max(0, (maxlen + 0) // 4 if maxlen > 3 else 0) or True
            acc = np.zeros(32)
            first = ((words[:, 1] & 0xffff) >> 3).astype(np.int64) if words.shape[1] > 1 else np.zeros(32, np.int64)
            # the kernel applies maxlen term steps (EXACT_STEPS) or every word of every chunk (padding is a no-op)
            nterm = min(maxlen, words.shape[1] - 1) if _exact({"vdot": 0, "jvs": 1, "lu": 2}.get(op, 4)) else words.shape[1] - 1
            for k in range(nterm if op not in ("div",) else 0):
                w = words[:, 1 + k]
                hi = ((w >> 16) >> 3).astype(np.int64)
                lo = ((w & 0xffff) >> 3).astype(np.int64)
                if op in ("vdot", "jvs"):
                    acc += self.coefs[hi] * kw["src"][lo]
                elif op == "lu":
                    G = kw["G"]
                    acc += G[hi] * G[lo]
                elif op in ("fwd", "bwd"):
                    acc += kw["G"][hi] * kw["X"][lo]
            g = 1 << lg
            if g > 1:
                acc = acc.reshape(-1, g).sum(axis=1).repeat(g)
            wr = (fl & 1) == 1
            if op == "vdot":
                kw["out"][row[wr]] = acc[wr]
            elif op == "jvs":
                G = kw["G"]
                G[row[wr]] = -acc[wr] + np.where((fl[wr] & 2) != 0, kw["ghinv"], 0.0)
                iv = wr & ((fl & 4) != 0)
                G[row[iv]] = 1.0 / G[row[iv]]
            elif op == "lu":
                G = kw["G"]
                G[row[wr]] = G[row[wr]] - acc[wr]
                iv = wr & ((fl & 4) != 0)
                G[row[iv]] = 1.0 / G[row[iv]]
            elif op == "div":
                G = kw["G"]
                a = wr & (ln > 0)
                G[row[a]] = (G[row[a]] * G[first[a]]) if self.recip_diag else (G[row[a]] / G[first[a]])
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
        G = np.zeros(self.mech.lu_nonzero + 1)        # + the zero slot
        for r in range(*self.phase["jvs"]):
            self.run_round(r, "jvs", src=B, G=G, ghinv=ghinv)
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
        """in place: L unit-lower multipliers, diagonal replaced by its reciprocal, U rows scaled"""
        for r in range(*self.phase["lu"]):
            self.run_round(r, "div" if self.rounds[r][2] & K_DIV else "lu", G=G)
        # tail: dense right-looking LU of the Schur complement, lane i = row i (kernel: registers + shuffles)
        m = self.m
        D = self._tail_dense(G)
        for j in range(m - 1):
            l = np.where(np.arange(m) > j, D[:, j] / D[j, j], 0.0)
            D[:, j] = np.where(np.arange(m) > j, l, D[:, j])
            for c in range(j + 1, m):
                D[:, c] = D[:, c] - l * D[j, c]
        for j in range(m):
            for i in range(m):
                p = self.tposT[j, i]
                if p != NONE:
                    G[p] = D[i, j]
                else:
                    assert D[i, j] == 0.0
        if self.recip_diag:
            G[self.diag[self.h:]] = 1.0 / G[self.diag[self.h:]]       # the head diagonals are reciprocals already
        else:
            G[self.diag] = 1.0 / G[self.diag]
        for i in range(self.n):
            G[self.diag[i] + 1:self.crow[i + 1]] *= G[self.diag[i]]
        return G

    def emulate_solve(self, G, X):
        m, h = self.m, self.h
        for r in range(*self.phase["fwd"]):
            self.run_round(r, "fwd", G=G, X=X)
        D = self._tail_dense(G)
        x = X[h:].copy()
        for j in range(m - 1):
            x = x - np.where(np.arange(m) > j, D[:, j], 0.0) * x[j]
        X[h:] = x
        X *= G[self.diag]
        x = X[h:].copy()
        for j in range(m - 1, 0, -1):
            x = x - np.where(np.arange(m) < j, D[:, j], 0.0) * x[j]
        X[h:] = x
        for r in range(*self.phase["bwd"]):
            self.run_round(r, "bwd", G=G, X=X)
        return X

    def stats(self):
        out = {}
        for name, (r0, r1) in self.phase.items():
            nb = [self.rounds[r][1] - self.rounds[r][0] for r in range(r0, r1)]
            rows = sum(bundle_chunks(self.bundles[b]).shape[0] for r in range(r0, r1) for b in range(self.rounds[r][0], self.rounds[r][1]))
            out[name] = dict(rounds=r1 - r0, bundles=sum(nb), bundles_per_round=nb, table_bytes=rows * 512)
        return out


if __name__ == "__main__":
    import sys
    from . import ir as IR
    mm = IR.load(sys.argv[1] if len(sys.argv) > 1 else "fullchem")
    s = Schedule(mm)
    for k, v in s.stats().items():
        print(k, v)
    print("bundles", len(s.bundles), "rounds", len(s.rounds), "coefs", s.coefs.size, "h", s.h, "m", s.m)
