"""Lane-level executable MODEL of the CUDA SCL kernel's bookkeeping (echoseal_b200/csrc/scl.cu):
thread-per-path, pointer-indirected alpha slots with no refcounts, bit-packed partial sums
(`bs` register for levels 6..10, pointer-indirected words for levels 1..5, in-place doubling
chain), candidate ranking with the reference's stable tie-break.  Runs on the CPU in pure
Python so the design can be checked against the oracle without a GPU
(tests/test_scl_lane_model.py).  Not product code, not the oracle.

The kernel has since moved the bottom two tree levels into registers (one "quad" of four decisions per
trip, partial sums combined once per quad) — a re-grouping of exactly these steps; the slot-pointer,
clone and ranking rules modelled here are unchanged."""
from __future__ import annotations
import math
import numpy as np

LOGE2 = 0.693147180559945309417232121458176568
NL = 8  # lanes per codeword


def phi(d):
    return math.log1p(math.exp(-d))


def logaddexp(x, y):
    if x == y:
        return x + LOGE2
    d = x - y
    mx = x if d > 0 else y
    return mx + phi(abs(d))


def fcomb(a, b):
    return logaddexp(a, b) - logaddexp(0.0, a + b)


def boff(lv):  # word offset of level lv (1..5) in the beta word store
    return (1 << (5 - lv)) - 1


def decode_lanes(llr, frozen, list_size=8):
    """Returns (paths sorted by metric: list of (metric, xhat_bits[1024])), following the kernel."""
    llr = [float(x) for x in llr]
    # alpha[lv][k][slot], lv = 1..9 ; leaf in lane register
    alpha = {lv: np.zeros((1 << (10 - lv), NL)) for lv in range(1, 10)}
    beta = np.zeros((31, NL), dtype=np.uint64)      # words for levels 1..5
    xroot = np.zeros((32, NL), dtype=np.uint64)
    active = [p == 0 for p in range(NL)]
    m = [0.0] * NL
    leaf = [0.0] * NL
    ptr = [0] * NL      # 3 bits per level 1..9
    bptr = [0] * NL     # 3 bits per level 1..5
    bs = [0] * NL       # packed small-level left partial sums
    order = [0] * NL

    def getp(word, lv):
        return (word >> (3 * (lv - 1))) & 7

    def setp(word, lv, v):
        return (word & ~(7 << (3 * (lv - 1)))) | (v << (3 * (lv - 1)))

    for i in range(1024):
        # ---------------- LLR update ----------------
        if i == 0:
            for lv in range(1, 11):          # cooperative spine, all into slot 0
                s = 1 << (10 - lv)
                for k in range(s):
                    a = llr[k] if lv == 1 else alpha[lv - 1][k, 0]
                    b = llr[k + s] if lv == 1 else alpha[lv - 1][k + s, 0]
                    v = fcomb(a, b)
                    if lv == 10:
                        leaf[0] = v
                    else:
                        alpha[lv][k, 0] = v
        else:
            t = (i & -i).bit_length() - 1
            l0 = 10 - t
            s = 1 << t
            new_vals = []
            for p in range(NL):              # g at level l0 (lanes in lockstep: reads before writes)
                ps = getp(ptr[p], l0 - 1) if l0 > 1 else 0
                out = []
                for k in range(s):
                    a = llr[k] if l0 == 1 else alpha[l0 - 1][k, ps]
                    b = llr[k + s] if l0 == 1 else alpha[l0 - 1][k + s, ps]
                    if s <= 16:
                        bit = (bs[p] >> (s + k)) & 1
                    else:
                        bsl = getp(bptr[p], l0)
                        bit = (int(beta[boff(l0) + (k >> 5), bsl]) >> (k & 31)) & 1
                    out.append(b - a if bit else b + a)
                new_vals.append(out)
            for p in range(NL):
                if l0 == 10:
                    leaf[p] = new_vals[p][0]
                else:
                    alpha[l0][:, p] = new_vals[p]
                    ptr[p] = setp(ptr[p], l0, p)
            for lv in range(l0 + 1, 11):     # f below, own slots
                s2 = 1 << (10 - lv)
                for p in range(NL):
                    for k in range(s2):
                        v = fcomb(alpha[lv - 1][k, p], alpha[lv - 1][k + s2, p])
                        if lv == 10:
                            leaf[p] = v
                        else:
                            alpha[lv][k, p] = v
                    if lv < 10:
                        ptr[p] = setp(ptr[p], lv, p)
        # ---------------- decision ----------------
        bit = [0] * NL
        if frozen[i]:
            for p in range(NL):
                if active[p]:
                    al = abs(leaf[p])
                    pen = phi(al) + (al if leaf[p] >= 0.0 else 0.0)
                    m[p] += pen
        else:
            m0 = [0.0] * NL; m1 = [0.0] * NL
            for p in range(NL):
                al = abs(leaf[p]); ph = phi(al)
                if leaf[p] >= 0.0:
                    m0[p] = m[p] + (ph + al); m1[p] = m[p] + ph
                else:
                    m0[p] = m[p] + ph; m1[p] = m[p] + (ph + al)
            r0 = [0] * NL; r1 = [0] * NL
            for p in range(NL):
                for j in range(NL):
                    if not active[j]:
                        continue
                    oj, o = order[j], order[p]
                    r0[p] += int(m0[j] < m0[p]) + int(m1[j] < m0[p]) + int(m0[j] == m0[p] and oj < o) + int(m1[j] == m0[p] and oj < o)
                    r1[p] += int(m0[j] < m1[p]) + int(m1[j] < m1[p]) + int(m0[j] == m1[p] and oj <= o) + int(m1[j] == m1[p] and oj < o)
            s0 = [active[p] and r0[p] < list_size for p in range(NL)]
            s1 = [active[p] and r1[p] < list_size for p in range(NL)]
            cm = [p for p in range(NL) if s0[p] and s1[p]]
            fm = [p for p in range(NL) if not (s0[p] or s1[p])]
            snap = [(m1[p], r1[p], ptr[p], bptr[p], bs[p]) for p in range(NL)]
            for p in range(NL):
                if s0[p]:
                    bit[p], m[p], order[p] = 0, m0[p], r0[p]
                elif s1[p]:
                    bit[p], m[p], order[p] = 1, m1[p], r1[p]
                else:
                    active[p] = False
            for j, f in enumerate(fm):
                if j < len(cm):
                    src = cm[j]
                    m[f], order[f], ptr[f], bptr[f], bs[f] = snap[src]
                    bit[f] = 1
                    active[f] = True
        # ---------------- partial sums ----------------
        newbeta = beta.copy(); newx = xroot.copy()
        for p in range(NL):
            b = bit[p]
            if (i & 1) == 0:
                bs[p] = (bs[p] & ~2) | (b << 1)
                continue
            t1 = ((~i) & (i + 1)).bit_length() - 1     # trailing ones
            X = b
            for j in range(min(t1, 5)):
                s = 1 << j
                Lb = (bs[p] >> s) & ((1 << s) - 1)
                X = (Lb ^ X) | (X << s)
            if t1 <= 4:
                s = 1 << t1
                bs[p] = (bs[p] & ~(((1 << s) - 1) << s)) | (X << s)
            elif t1 == 5:
                newbeta[boff(5), p] = X
                bptr[p] = setp(bptr[p], 5, p)
            else:
                lstar = 10 - t1
                D = [X]
                for l in range(5, lstar, -1):
                    bsl = getp(bptr[p], l)
                    n = len(D)
                    Lw = [int(beta[boff(l) + w, bsl]) for w in range(n)]
                    D = [D[w] ^ Lw[w] for w in range(n)] + D
                if lstar > 0:
                    for w, v in enumerate(D):
                        newbeta[boff(lstar) + w, p] = v
                    bptr[p] = setp(bptr[p], lstar, p)
                else:
                    for w, v in enumerate(D):
                        newx[w, p] = v
        beta, xroot = newbeta, newx
    res = []
    for p in range(NL):
        if not active[p]:
            continue
        x = np.zeros(1024, np.uint8)
        for w in range(32):
            for bpos in range(32):
                x[32 * w + bpos] = (int(xroot[w, p]) >> bpos) & 1
        res.append((m[p], order[p], x))
    res.sort(key=lambda r: (r[0], r[1]))
    return [(r[0], r[2]) for r in res]
