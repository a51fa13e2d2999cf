"""Lane-level simulation (numpy) of the register-resident Jacobi proposed for ensi_kernel in DESIGN.md section 9:
lane i holds row i of A and of V in registers, the round-robin rounds are unrolled (static column indices), the row
phase exchanges whole rows between the two lanes of a pair with shuffles. The simulation uses exactly the index formulas
the kernel would use and checks the result against numpy.linalg.eigh, cold and warm-started.

  partner(r, i) = r                 if i == M          (M = E - 1, E even)
                = M                 if i == r
                = (2 r - i) mod M   otherwise          (the two indices of a pair sum to 2 r mod M)
  slot(r, i)    = 0 if i in (M, r) else min(l, M - l), l = (i - r) mod M
  pair of slot t in round r: t == 0 -> (r, M); else ((r + t) mod M, (r - t) mod M)

usage: python profiles/jacobi_systolic_sim.py
"""
import numpy as np


def partner(r, i, M):
    if i == M:
        return r
    if i == r:
        return M
    return (2 * r - i) % M


def slot(r, i, M):
    if i == M or i == r:
        return 0
    l = (i - r) % M
    return min(l, M - l)


def pair_of_slot(r, t, M):
    p, q = (r, M) if t == 0 else ((r + t) % M, (r - t) % M)
    return (p, q) if p < q else (q, p)


def sweep(a, v, E):
    """One sweep of M rounds on the lane-resident rows a[i, :], v[i, :]. Returns the number of rotations applied."""
    M, H = E - 1, E // 2
    rotations = 0
    for r in range(M):
        # every lane finds its pair; both lanes of a pair compute the same angle from (app, aqq, apq)
        cs = np.zeros((H, 2))
        cs[:, 0] = 1.0
        for i in range(E):
            j = partner(r, i, M)
            assert partner(r, j, M) == i and slot(r, i, M) == slot(r, j, M)
            p, q = min(i, j), max(i, j)
            assert pair_of_slot(r, slot(r, i, M), M) == (p, q)
            if i != p:
                continue
            app, aqq, apq = a[p, p], a[q, q], a[p, q]
            if apq * apq > 1e-30 * abs(app * aqq):
                theta = (aqq - app) / (2 * apq)
                t = np.sign(theta) / (abs(theta) + np.sqrt(theta * theta + 1)) if theta != 0 else 1.0
                c = 1 / np.sqrt(t * t + 1)
                cs[slot(r, i, M)] = (c, t * c)
                rotations += 1
        # column phase: static (p_t, q_t) per slot, every lane rotates its own two elements of A and of V
        for t in range(H):
            p, q = pair_of_slot(r, t, M)
            c, s = cs[t]
            for m in (a, v):
                mp, mq = m[:, p].copy(), m[:, q].copy()
                m[:, p] = c * mp - s * mq
                m[:, q] = s * mp + c * mq
        # row phase (A only): lane i gets its partner's row by shuffle; new_p = c row_p - s row_q, new_q = s row_p + c row_q
        old = a.copy()          # the shuffles read the rows as they are after the column phase
        for i in range(E):
            j = partner(r, i, M)
            c, s = cs[slot(r, i, M)]
            ss = -s if i < j else s
            a[i, :] = c * old[i, :] + ss * old[j, :]
    return rotations


def decompose(A0, V0=None, E=20):
    v = np.eye(E) if V0 is None else V0.copy()
    a = v.T @ A0 @ v          # warm start: A = V' Pinv V (the identity for a cold start)
    sweeps = 0
    while True:
        off = (a * a).sum() - (np.diag(a) ** 2).sum()
        if off <= 1e-26 * (np.diag(a) ** 2).sum():
            break
        sweep(a, v, E)
        sweeps += 1
        assert sweeps < 30
    return np.diag(a).copy(), v, sweeps


def main():
    rng = np.random.default_rng(3)
    E, k = 20, 50
    Y = rng.normal(size=(k, E))
    rinv = rng.uniform(0.1, 4.0, k)
    A0 = Y.T @ (rinv[:, None] * Y) + (E - 1) * np.eye(E)
    lam, V, n_cold = decompose(A0, E=E)
    want = np.linalg.eigvalsh(A0)
    assert np.allclose(np.sort(lam), want, rtol=1e-12), "eigenvalues differ"
    assert np.allclose(V @ np.diag(lam) @ V.T, A0, rtol=1e-11, atol=1e-11) and np.allclose(V.T @ V, np.eye(E), atol=1e-12)
    # a neighbouring point: slightly different weights, warm-started from V
    A1 = Y.T @ ((rinv * (1 + 0.01 * rng.normal(size=k)))[:, None] * Y) + (E - 1) * np.eye(E)
    lam1, V1, n_warm = decompose(A1, V0=V, E=E)
    assert np.allclose(np.sort(lam1), np.linalg.eigvalsh(A1), rtol=1e-12)
    W = V1 @ np.diag(np.sqrt((E - 1) / lam1)) @ V1.T
    w_ref, v_ref = np.linalg.eigh(A1)
    assert np.allclose(W, v_ref @ np.diag(np.sqrt((E - 1) / w_ref)) @ v_ref.T, rtol=1e-10, atol=1e-12)
    print("ok: cold start %d sweeps, warm start %d sweeps; schedule formulas consistent for E = %d" % (n_cold, n_warm, E))
    for E2 in (10, 30):
        Y2 = rng.normal(size=(40, E2))
        A2 = Y2.T @ Y2 + (E2 - 1) * np.eye(E2)
        lam2, _, n2 = decompose(A2, E=E2)
        assert np.allclose(np.sort(lam2), np.linalg.eigvalsh(A2), rtol=1e-12)
        print("ok: E = %d, %d sweeps" % (E2, n2))


if __name__ == "__main__":
    main()
