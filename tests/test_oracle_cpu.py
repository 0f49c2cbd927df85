"""CPU tests: pin the oracle (oracle/) against the reference's golden vectors and against the
third-party routines the reference calls (numpy reductions, scipy.linalg.qr)."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import clib, pod_oracle as po, synth


# ---- oracle == unmodified reference on the committed fixtures (oracle/make_golden.py) ----------
def test_oracle_matches_reference_golden(golden):
    g = golden
    X = g["X"]
    n_c = X.shape[0] // g["F"]
    f = po.fit(X, g["F"], g["scale_type"], g["axis_cnt"], g["select_modes"], g["n_modes"])
    np.testing.assert_array_equal(f["X_cnt"], g["X_cnt"])
    np.testing.assert_array_equal(f["X_scl"], g["X_scl"])
    assert f["r"] == g["r"]
    np.testing.assert_array_equal(f["Ur"], g["Ur"])
    np.testing.assert_array_equal(f["Ar"], g["Ar"])
    np.testing.assert_array_equal(f["Sigma_r"], g["Sigma_r"])
    np.testing.assert_array_equal(f["X0"][:64], g["X0_head"])
    piv = po.qr_pivots(f["Ur"])
    np.testing.assert_array_equal(piv, g["piv"])
    C = po.one_hot(piv, X.shape[0])
    Th = po.theta(C, f["Ur"])
    np.testing.assert_array_equal(Th, g["Theta"])
    Ar, Asig = po.predict_ols(Th, list(g["Y"]), C, f["X_cnt"], f["X_scl"], n_c)
    np.testing.assert_array_equal(Ar, g["Ar_pred"])
    np.testing.assert_array_equal(Asig, g["Ar_sigma"])
    np.testing.assert_array_equal(po.reconstruct(f["Ur"], Ar, f["X_cnt"], f["X_scl"]), g["X_rec"])


# ---- C restatement of numpy's pairwise tree == numpy ------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 15, 16, 41, 127, 128, 129, 130, 255, 256, 257, 1000,
                               1024, 4097, 10007, 65536, 100003])
def test_pairwise_sum_bit_exact(n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
    assert clib.pairwise_sum(a) == np.add.reduce(a)


@pytest.mark.parametrize("shape", [(10, 5), (400, 41), (300, 160), (50, 1024), (2000, 41), (77, 129)])
def test_block_stats_and_row_means_bit_exact(shape):
    rng = np.random.default_rng(shape[0])
    x = rng.random(shape) * 3.0 + 1.0
    s, mu, q, lo, hi = clib.block_stats(x)
    assert mu == np.average(x)
    assert np.sqrt(q / x.size) == np.std(x)
    assert q / x.size == np.var(x)
    assert (lo, hi) == (np.min(x), np.max(x))
    np.testing.assert_array_equal(clib.row_means(x), np.average(x, axis=1))


# ---- C restatement of dlaqp2 == scipy.linalg.qr(pivoting=True) --------------------------------
def _orth(n, r, seed):
    rng = np.random.default_rng(seed)
    w = 1.0 + 5.0 * rng.random((n, 1)) ** 4
    Q, _ = np.linalg.qr(rng.standard_normal((n, r)) * w)
    return Q


@pytest.mark.parametrize("n,r", [(50, 5), (2000, 14), (5000, 40), (6000, 100), (3000, 128)])
def test_dlaqp2_matches_scipy(n, r):
    Ur = _orth(n, r, 7 * n + r)
    _, R, P = sla.qr(Ur.T, pivoting=True, mode="economic")
    o = clib.qrcp_dlaqp2(Ur)
    np.testing.assert_array_equal(o["piv"], P[:r])
    np.testing.assert_allclose(np.abs(o["rdiag"]), np.abs(np.diag(R)), rtol=1e-12)
    assert o["gap"].min() > 1e-9


def test_dlaqp2_matches_reference_golden(golden):
    o = clib.qrcp_dlaqp2(golden["Ur"])
    np.testing.assert_array_equal(o["piv"], golden["piv"])


def test_dlaqp2_exact_ties_follow_lapack_swaps():
    """Duplicated columns tie at every step; the winner is decided by LAPACK's permuted order."""
    for seed in range(12):
        rng = np.random.default_rng(seed)
        base = _orth(60, 6, seed)
        Ur = np.concatenate([base, base[rng.permutation(60)[:30]], base], axis=0)
        _, _, P = sla.qr(Ur.T, pivoting=True, mode="economic")
        np.testing.assert_array_equal(clib.qrcp_dlaqp2(Ur)["piv"], P[:6])


def test_dlaqp2_masked_rows():
    Ur = _orth(800, 10, 3)
    mask = np.ones(800, dtype=bool)
    mask[100:500] = False
    piv = po.qr_pivots(Ur, mask)
    Um = Ur.copy()
    Um[~mask] = 0
    np.testing.assert_array_equal(clib.qrcp_dlaqp2(Um)["piv"], piv)
    assert mask[piv].all()


# ---- synthetic generator: shardable and deterministic -----------------------------------------
def test_synth_shards_reassemble_bit_exact():
    F, n_c, m, r = 3, 50, 17, 6
    full = synth.snapshots(F, n_c, m, r)
    parts = [synth.snapshots(F, n_c, m, r, cell0=c0, ncell_loc=nl) for c0, nl in ((0, 20), (20, 30))]
    for f in range(F):
        got = np.concatenate([p[f * p.shape[0] // F:(f + 1) * p.shape[0] // F] for p in parts])
        np.testing.assert_array_equal(got, full[f * n_c:(f + 1) * n_c])
    assert np.all(full > 0)


def test_synth_spectrum_is_benign_for_gram_path():
    X = synth.snapshots(9, 2000, 41, 40)
    f = po.fit(X, 9, n_modes=40, select_modes="number")
    s = f["Sigma_r"]
    assert s[-1] / s[0] > 1e-3          # Gram-eigh error eps*(s1/sr)^2 stays below 1e-10
    assert np.min(s[:-1] / s[1:]) > 1.02  # distinct singular values -> modes defined up to sign
    assert clib.qrcp_dlaqp2(f["Ur"])["gap"].min() > 1e-7


# ---- reference error behaviour restated by the oracle (sparse_sensing.py:69-81, :314-333) -----
def test_oracle_validation_errors():
    X = np.zeros((6, 3))
    with pytest.raises(TypeError):
        po.check_inputs([[1.0]], 1)
    with pytest.raises(TypeError):
        po.check_inputs(X, 2.0)
    with pytest.raises(Exception):
        po.check_inputs(X, 4)
    ev = np.array([50.0, 90.0, 100.0])
    assert po.choose_rank(ev, 3, "variance", 80) == 2
    assert po.choose_rank(ev, 3, "variance", 100) == 3
    with pytest.raises(ValueError):
        po.choose_rank(ev, 3, "variance", 101)
    with pytest.raises(TypeError):
        po.choose_rank(ev, 3, "number", 2.0)
    with pytest.raises(ValueError):
        po.choose_rank(ev, 3, "number", 4)
    with pytest.raises(ValueError):
        po.choose_rank(ev, 3, "bogus", 1)


# ---- GEM placement: oracle == unmodified reference (sparse_sensing.py:586-698), jitter reproduced
#      by seeding numpy's global generator exactly like oracle/make_golden.py did -----------------
def test_gem_oracle_matches_reference_golden(golden_gem):
    g = golden_gem
    np.random.seed(int(g["seed"]))
    sensors, cond = po.gem_placement(g["Ur"], g["xyz"], int(g["F"]), int(g["n_sensors"]), g["mask"], float(g["d_min"]))
    np.testing.assert_array_equal(sensors, g["gem"])
    assert np.all(g["mask"][sensors])
    if float(g["d_min"]) > 0:
        P = np.tile(g["xyz"], (int(g["F"]), 1))[sensors]
        D = np.linalg.norm(P[:, None, :] - P[None, :, :], axis=2)
        assert np.all(D[np.triu_indices(len(sensors), 1)] >= float(g["d_min"]))


# ---- every scaling type x both centring modes: oracle == unmodified reference, bit for bit -------
SCALINGS = ("std", "none", "pareto", "vast", "range", "level", "max", "variance", "median", "poisson", "l2-norm")


@pytest.mark.parametrize("scale_type", SCALINGS)
@pytest.mark.parametrize("axis_cnt,tag", [(1, "row"), (None, "blk")])
def test_oracle_scalings_match_reference_golden(scale_type, axis_cnt, tag):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "s1_scalings_450x11.npz"))
    X0, cnt, scl = po.center_scale(z["X"], int(z["F"]), scale_type, axis_cnt)
    key = scale_type.replace("-", "_") + "_" + tag
    np.testing.assert_array_equal(cnt, z[key + "_cnt"])
    np.testing.assert_array_equal(scl, z[key + "_scl"])
    np.testing.assert_array_equal(X0, z[key + "_X0"])
