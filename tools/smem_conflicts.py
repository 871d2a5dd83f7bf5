#!/usr/bin/env python
"""Count shared-memory wavefronts of the pencil kernel's inter-stage writes for a choice of
slab-stride residues (8-byte accesses: a half-warp of 16 lanes is conflict-free iff its 16
addresses fall in 16 distinct 8-byte banks).  Used to pick the pad residues in pa_apply.cuh."""
import itertools
import sys


def wavefronts(addrs):
    """addrs: list of per-lane double indices for one warp instruction (None = inactive)."""
    w = 0
    for h in range(0, len(addrs), 16):
        lanes = [a for a in addrs[h:h + 16] if a is not None]
        if not lanes:
            continue
        banks = {}
        for a in lanes:
            banks.setdefault(a % 16, set()).add(a)
        w += max(len(v) for v in banks.values())
    return w


def stage_writes(D, Q, NEB, rA, rC, rE, rF):
    """total wavefronts and ideal for the four write patterns, one field each"""
    T1, T2, T3 = NEB * D * D, NEB * Q * D, NEB * Q * Q
    def pad(n, r):
        m = n - n % 16 + r % 16
        return m if m >= n else m + 16
    SA, SC, SE, SF = pad(T2, rA), pad(T3, rC), pad(T2, rE), pad(T1, rF)
    NT = (T3 + 31) // 32 * 32
    res = {}
    # A: thread t1=(dx,dy,e), fixed qz
    tot = ideal = 0
    for qz in range(Q):
        for w0 in range(0, NT, 32):
            ad = []
            for t in range(w0, w0 + 32):
                if t < T1:
                    e, r = divmod(t, D * D); dy, dx = divmod(r, D)
                    ad.append(dy * SA + dx + D * (qz + Q * e))
                else:
                    ad.append(None)
            tot += wavefronts(ad); ideal += (sum(a is not None for a in ad) + 15) // 16
    res["A"] = (tot, ideal)
    tot = ideal = 0
    for qy in range(Q):
        for w0 in range(0, NT, 32):
            ad = []
            for t in range(w0, w0 + 32):
                if t < T2:
                    zq, dx = divmod(t, D)
                    ad.append(dx * SC + Q * zq + qy)
                else:
                    ad.append(None)
            tot += wavefronts(ad); ideal += (sum(a is not None for a in ad) + 15) // 16
    res["C"] = (tot, ideal)
    tot = ideal = 0
    for dx in range(D):
        for w0 in range(0, NT, 32):
            ad = []
            for t in range(w0, w0 + 32):
                if t < T3:
                    zq, qy = divmod(t, Q)
                    ad.append(qy * SE + D * zq + dx)
                else:
                    ad.append(None)
            tot += wavefronts(ad); ideal += (sum(a is not None for a in ad) + 15) // 16
    res["E"] = (tot, ideal)
    tot = ideal = 0
    for dy in range(D):
        for w0 in range(0, NT, 32):
            ad = []
            for t in range(w0, w0 + 32):
                if t < T2:
                    zq, dx = divmod(t, D); e, qz = divmod(zq, Q)
                    ad.append(qz * SF + dx + D * dy + D * D * e)
                else:
                    ad.append(None)
            tot += wavefronts(ad); ideal += (sum(a is not None for a in ad) + 15) // 16
    res["F"] = (tot, ideal)
    return res


if __name__ == "__main__":
    D, Q, NEB = (int(v) for v in sys.argv[1:4])
    cur = stage_writes(D, Q, NEB, D, 1, 1, D)
    print("current residues (D,1,1,D):", cur)
    best = {}
    for name, idx in (("A", 0), ("C", 1), ("E", 2), ("F", 3)):
        cands = []
        for r in range(16):
            rs = [D, 1, 1, D]
            rs[idx] = r
            t, i = stage_writes(D, Q, NEB, *rs)[name]
            cands.append((t, r, i))
        cands.sort()
        best[name] = cands[:3]
    print("best residues per slab (wavefronts, residue, ideal):", best)


def search_rowpad(D, Q, NEB):
    """generalised layouts: row stride R (>= natural) inside a slab, slab stride residue s."""
    T1, T2, T3 = NEB * D * D, NEB * Q * D, NEB * Q * Q
    NT = (T3 + 31) // 32 * 32
    ZQ = Q * NEB

    def total(gen_w, gen_r, nT_w, nT_r, loops_w, loops_r):
        tw = tr = 0
        for l in range(loops_w):
            for w0 in range(0, NT, 32):
                tw += wavefronts([gen_w(t, l) if t < nT_w else None for t in range(w0, w0 + 32)])
        for l in range(loops_r):
            for w0 in range(0, NT, 32):
                tr += wavefronts([gen_r(t, l) if t < nT_r else None for t in range(w0, w0 + 32)])
        return tw, tr

    out = {}
    # C: W thread t2=(dx,zq) loop qy ; R thread t3=(qy,zq) loop dx ; addr = dx*S + zq*R + qy
    best = []
    for R in range(Q, Q + 4):
        for s in range(16):
            S = ZQ * R; S = S - S % 16 + s; S = S if S >= ZQ * R else S + 16
            tw, tr = total(lambda t, qy: (t % D) * S + (t // D) * R + qy,
                           lambda t, dx: dx * S + (t // Q) * R + (t % Q), T2, T3, Q, D)
            best.append((tw + tr, tw, tr, R, s, S))
    out["C"] = sorted(best)[:4]
    # E: W thread t3=(qy,zq) loop dx ; R thread t2=(dx,zq) loop qy ; addr = qy*S + zq*R + dx
    best = []
    for R in range(D, D + 4):
        for s in range(16):
            S = ZQ * R; S = S - S % 16 + s; S = S if S >= ZQ * R else S + 16
            tw, tr = total(lambda t, dx: (t % Q) * S + (t // Q) * R + dx,
                           lambda t, qy: qy * S + (t // D) * R + (t % D), T3, T2, D, Q)
            best.append((tw + tr, tw, tr, R, s, S))
    out["E"] = sorted(best)[:4]
    # A: W thread t1=(dx,dy,e) loop qz ; R thread t2=(dx,zq) loop dy ; addr = dy*S + zq*R + dx
    best = []
    for R in range(D, D + 4):
        for s in range(16):
            S = ZQ * R; S = S - S % 16 + s; S = S if S >= ZQ * R else S + 16
            tw, tr = total(lambda t, qz: ((t % (D * D)) // D) * S + (qz + Q * (t // (D * D))) * R + t % D,
                           lambda t, dy: dy * S + (t // D) * R + (t % D), T1, T2, Q, D)
            best.append((tw + tr, tw, tr, R, s, S))
    out["A"] = sorted(best)[:4]
    # F: W thread t2=(dx,qz,e) loop dy ; R thread t1=(dx,dy,e) loop qz ; addr = qz*S + (dy + D*e)*R + dx
    best = []
    DE = D * NEB
    for R in range(D, D + 4):
        for s in range(16):
            S = DE * R; S = S - S % 16 + s; S = S if S >= DE * R else S + 16
            tw, tr = total(lambda t, dy: ((t // D) % Q) * S + (dy + D * (t // (D * Q))) * R + t % D,
                           lambda t, qz: qz * S + (((t % (D * D)) // D) + D * (t // (D * D))) * R + t % D,
                           T2, T1, D, Q)
            best.append((tw + tr, tw, tr, R, s, S))
    out["F"] = sorted(best)[:4]
    ideal = {"C": ((T2 + 15) // 16 * Q, (T3 + 15) // 16 * D), "E": ((T3 + 15) // 16 * D, (T2 + 15) // 16 * Q),
             "A": ((T1 + 15) // 16 * Q, (T2 + 15) // 16 * D), "F": ((T2 + 15) // 16 * D, (T1 + 15) // 16 * Q)}
    return out, ideal


if __name__ == "__main__" and len(sys.argv) > 4:
    o, ideal = search_rowpad(D, Q, NEB)
    for k in o:
        print(k, "ideal (w,r)", ideal[k], "best (tot, w, r, R, s, S):", o[k])
