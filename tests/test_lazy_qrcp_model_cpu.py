"""CPU model of the lazy norm down-dates of the blocked pivoted QR (openmeasure_b200/csrc/qrcp.cu, file
header): the same decisions in numpy -- segments of 64 candidates, the bound theta = alpha * (pivot norm at
the block start), certification of every in-block pivot, catch-up of the segments that could beat an
uncertified one, exact norms at every block boundary -- checked against the eager schedule and against
LAPACK (scipy.linalg.qr(Ur.T, pivoting=True), the call the reference makes at sparse_sensing.py:739).

This does not exercise the CUDA kernels (tests/test_gpu_parity.py does); it pins down that the SCHEME is
exact: whatever alpha, the pivots are those of the eager schedule, and a segment is never read before it has to be."""
import numpy as np
import pytest
import scipy.linalg as sla

SEG = 64
TOL3Z = np.sqrt(np.finfo(np.float64).eps)


def _downdate(v1, v2, rij, tail_norm):
    """dlaqp2's partial-norm down-date of one column (LAPACK working note 176)."""
    if v1 <= 0.0:
        return v1, v2
    temp = max(1.0 - (abs(rij) / v1) ** 2, 0.0)
    temp2 = temp * (v1 / v2) ** 2
    if temp2 <= TOL3Z:
        v = tail_norm()
        return v, v
    return v1 * np.sqrt(temp), v2


def blocked_qrcp(A, block, alpha, s=None):
    """Blocked QRCP of the r x n matrix A (= Ur^T).  alpha = 0: every pass visits every segment.
    Returns (pivots, visits) with visits = (segment, row) reads of the in-block passes."""
    A = A.copy()
    r, n = A.shape
    s = r if s is None else s
    nseg = -(-n // SEG)
    seg_of = np.arange(n) // SEG
    vn1 = np.linalg.norm(A, axis=0)
    vn2 = vn1.copy()
    alive = np.ones(n, bool)
    piv, visits = [], 0
    i0 = 0
    while len(piv) < s:
        L = r - i0
        A0 = A[i0:, :].copy()                     # trailing matrix at the block start (never written inside the block)
        seg_max = np.array([np.max(np.where(alive[k * SEG:(k + 1) * SEG], vn1[k * SEG:(k + 1) * SEG], -1.0), initial=-1.0)
                            for k in range(nseg)])
        V, tau, Q = [], [], np.eye(L)
        skip = np.zeros(nseg, bool)
        theta = -1.0

        def bring(cols, t_lo, t_hi):
            """down-dates of steps t_lo .. t_hi - 1 for the given columns (R[t, j] = q_t . a_j)."""
            nonlocal visits
            for t in range(t_lo, t_hi):
                q = Q[:, t]
                for j in cols:
                    if not alive[j]:
                        continue
                    rij = q @ A0[:, j]
                    vn1[j], vn2[j] = _downdate(vn1[j], vn2[j], rij,
                                               lambda j=j, t=t: np.linalg.norm((Q.T @ A0[:, j])[t + 1:]))
            visits += len(set(seg_of[cols])) * L * (t_hi - t_lo) if len(cols) else 0

        for t in range(min(block, s - len(piv))):
            def best(mask):
                cand = np.where(mask & alive, vn1, -1.0)
                return int(np.argmax(cand)), float(cand.max())      # ties -> lowest index, as idamax on unswapped columns
            active_cols = ~skip[seg_of]
            p, c = best(active_cols)
            if t == 0:
                theta = alpha * c
                skip = seg_max < theta if alpha > 0 else np.zeros(nseg, bool)
            elif skip.any() and not c >= theta:
                # not certified: the skipped segments that could beat c catch up, theta drops to c
                join = skip & (seg_max >= c)
                bring(np.nonzero(join[seg_of])[0], 0, t)
                skip &= ~join
                theta = c
                p, c = best(~skip[seg_of])
            assert not skip.any() or np.all(seg_max[skip] < c)       # everything still skipped is below the pivot
            piv.append(p)
            alive[p] = False
            # reflector of step t from the pivot column brought up to date, q_t = Q e_t
            x = Q.T @ A0[:, p]
            beta = -np.copysign(np.linalg.norm(x[t:]), x[t])
            v = np.zeros(L)
            v[t:] = x[t:]
            v[t] -= beta
            tq = 0.0 if np.linalg.norm(v) == 0 else 2.0 / (v @ v)
            Q = Q - tq * np.outer(Q @ v, v)
            if len(piv) == s:
                break
            if t < block - 1:
                bring(np.nonzero(~skip[seg_of])[0], t, t + 1)        # the read-only pass: active segments only
        # block-closing pass: every column is updated and leaves with its EXACT trailing norm
        nb = t + 1
        A[i0:, :] = Q.T @ A0
        ex = np.linalg.norm(A[i0 + nb:, :], axis=0)
        vn1 = np.where(alive, ex, vn1)
        vn2 = vn1.copy()
        i0 += nb
    return np.array(piv), visits


def _orth(n, r, seed, localised=False):
    rng = np.random.default_rng(seed)
    if localised:
        x = np.linspace(0, 1, n)[:, None]
        A = np.exp(-((x - rng.random(r)[None, :]) / (0.03 + 0.2 * rng.random(r)[None, :])) ** 2)
        A = A * np.cos(2 * np.pi * x * (1 + np.arange(r))[None, :]) + 1e-3 * rng.standard_normal((n, r))
    else:
        A = rng.standard_normal((n, r)) * (1.0 + 5.0 * rng.random((n, 1)) ** 4)
    return np.linalg.qr(A)[0]


@pytest.mark.parametrize("n,r,block,localised", [(700, 12, 4, False), (1500, 20, 8, False), (1500, 20, 8, True),
                                                  (2000, 16, 3, True), (900, 9, 8, False)])
def test_lazy_schedule_picks_lapack_pivots(n, r, block, localised):
    Ur = _orth(n, r, 11 * n + r, localised)
    _, _, P = sla.qr(Ur.T, pivoting=True, mode="economic")
    eager, v0 = blocked_qrcp(Ur.T, block, 0.0)
    np.testing.assert_array_equal(eager, P[:r])
    for alpha in (0.5, 0.9, 0.94, 0.999):
        lazy, v = blocked_qrcp(Ur.T, block, alpha)
        np.testing.assert_array_equal(lazy, eager)
        assert v <= v0
    if localised:
        assert blocked_qrcp(Ur.T, block, 0.94)[1] < 0.5 * v0           # most of a localised mesh is never read in a block


def test_lazy_schedule_keeps_exact_ties_in_lapack_order():
    for seed in range(3):
        rng = np.random.default_rng(seed)
        base = _orth(200, 6, seed)
        Ur = np.concatenate([base, base[rng.permutation(200)[:130]], 0.0 * base[:70], base], axis=0)
        _, _, P = sla.qr(Ur.T, pivoting=True, mode="economic")
        for alpha in (0.0, 0.94, 0.999):
            piv, _ = blocked_qrcp(Ur.T, 4, alpha)
            # duplicated rows have identical norms whether or not their segment was visited: the first copy wins
            np.testing.assert_array_equal(np.sort(piv), np.sort(P[:6]))
            assert all(np.allclose(Ur[a], Ur[b]) for a, b in zip(piv, P[:6]))
