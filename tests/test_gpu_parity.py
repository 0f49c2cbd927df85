"""GPU parity tests: the CUDA path (through the C ABI / the ROM-SPR mirror) against the CPU oracle,
the reference's golden fixtures, and numpy/scipy themselves.  Tolerances follow BASELINE.json:
pivots bit-exact on non-degenerate inputs, singular values and reconstructions 1e-10 relative,
modes up to sign; centring/scaling bit-exact (the reference's own tests use assert_array_equal)."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu

RTOL = 1e-10


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    from openmeasure_b200 import build
    build.build()
    return torch


def _sps():
    from openmeasure_b200 import sparse_sensing
    return sparse_sensing


def _orth(n, r, seed):
    rng = np.random.default_rng(seed)
    w = 1.0 + 5.0 * rng.random((n, 1)) ** 4
    Q, _ = np.linalg.qr(rng.standard_normal((n, r)) * w)
    return Q


def _sign_align(U, Uref):
    s = np.sign(np.sum(U * Uref, axis=0))
    s[s == 0] = 1
    return U * s, s


# ---------------------------------------------------------------------------------------------
# synthetic generator: CUDA twin == CPU twin, bit for bit
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,n_c,m,r,c0,ncl", [(3, 50, 17, 6, 0, 50), (9, 700, 41, 14, 100, 333),
                                              (2, 300, 160, 20, 0, 300)])
def test_synth_cuda_equals_cpu(torch_cuda, F, n_c, m, r, c0, ncl):
    from openmeasure_b200 import synth as gsynth
    from oracle import synth as osynth
    Xg = gsynth.snapshots(F, n_c, m, r, cell0=c0, ncell_loc=ncl).cpu().numpy()
    Xo = osynth.snapshots(F, n_c, m, r, cell0=c0, ncell_loc=ncl)
    np.testing.assert_array_equal(Xg, Xo)


# ---------------------------------------------------------------------------------------------
# K1: centring / scaling, bit-exact vs numpy (mirrors reference tests/test_rom.py:19-46)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_pts,F,m", [(10, 2, 5), (1, 3, 4), (400, 3, 41), (77, 2, 129), (50, 2, 1024),
                                       (3000, 4, 41), (7, 5, 300), (20000, 2, 33)])
@pytest.mark.parametrize("scale_type", ["std", "pareto", "range"])
def test_centering_and_scaling_bit_exact(torch_cuda, n_pts, F, m, scale_type):
    rng = np.random.default_rng(n_pts * 31 + m)
    X = rng.random((n_pts * F, m)) * 10.0 ** rng.integers(-2, 3, (n_pts * F, 1))
    rom = _sps().ROM(X, F, None)
    X0 = rom.scale_data(scale_type)
    np.testing.assert_array_equal(rom.X_cnt, np.mean(X, axis=1)[:, np.newaxis])
    X_scl = np.zeros((X.shape[0], 1))
    for f in range(F):
        blk = X[f * n_pts:(f + 1) * n_pts]
        X_scl[f * n_pts:(f + 1) * n_pts] = {"std": np.std(blk), "pareto": np.sqrt(np.std(blk)),
                                            "range": np.max(blk) - np.min(blk)}[scale_type]
    np.testing.assert_array_equal(rom.X_scl, X_scl)
    np.testing.assert_array_equal(X0, (X - np.mean(X, axis=1)[:, np.newaxis]) / X_scl)


@pytest.mark.parametrize("scale_type", ["none", "vast", "level", "max", "variance", "poisson", "median",
                                        "l2-norm"])
def test_other_scalings_match_oracle(torch_cuda, scale_type):
    from oracle import pod_oracle as po
    rng = np.random.default_rng(5)
    X = rng.random((3 * 211, 19)) + 0.25
    rom = _sps().ROM(X, 3, None)
    X0 = rom.scale_data(scale_type)
    X0o, cnt, scl = po.center_scale(X, 3, scale_type, 1)
    np.testing.assert_array_equal(rom.X_cnt, cnt)
    np.testing.assert_allclose(rom.X_scl, scl, rtol=4e-16 if scale_type != "l2-norm" else 1e-13)
    np.testing.assert_allclose(X0, X0o, rtol=1e-13, atol=1e-15)


def test_centering_axis_none_bit_exact(torch_cuda):
    rng = np.random.default_rng(9)
    X = rng.random((2 * 1500, 23))
    rom = _sps().ROM(X, 2, None)
    rom.scale_data(axis_cnt=None)
    X_cnt = np.zeros((X.shape[0], 1))
    for f in range(2):
        X_cnt[f * 1500:(f + 1) * 1500] = np.mean(X[f * 1500:(f + 1) * 1500])
    np.testing.assert_array_equal(rom.X_cnt, X_cnt)


def test_broken_options_raise_like_reference(torch_cuda):
    X = np.random.default_rng(0).random((20, 5))
    rom = _sps().ROM(X, 2, None)
    with pytest.raises(ValueError):
        rom.scale_data("vast_2")
    with pytest.raises(ValueError):
        rom.scale_data(axis_cnt=0)
    with pytest.raises(NotImplementedError):
        rom.scale_data("nonsense")


# ---------------------------------------------------------------------------------------------
# K6: pivoted QR kernel vs the C oracle (bitwise, block=1), scipy and the blocked variant
# ---------------------------------------------------------------------------------------------
def _gpu_qrcp(torch, Ur, block, s=None, lazy=None, stats=None):
    """lazy: alpha of the lazy norm down-dates for this call (None: the library default; 0: eager)."""
    from openmeasure_b200 import _lib, engine
    n, r = Ur.shape
    eng = engine.Engine(torch.zeros(n, 1, dtype=torch.float64, device="cuda"), 1, group=False)
    eng.set_basis_rows(torch.from_numpy(np.ascontiguousarray(Ur)).cuda())
    prev = _lib.load().omb_qrcp_set_lazy(lazy) if lazy is not None else None
    try:
        piv, rdiag, gap = eng.qrcp(s=s, block=block)
        if stats is not None:
            stats.update(eng.qr_stats())
    finally:
        if prev is not None:
            _lib.load().omb_qrcp_set_lazy(prev)
    return piv.cpu().numpy(), rdiag.cpu().numpy(), gap.cpu().numpy()


@pytest.mark.parametrize("n,r", [(20, 5), (50, 5), (2001, 14), (5000, 40), (6000, 100), (3001, 128),
                                 (100000, 24)])
def test_qrcp_unblocked_is_bitwise_dlaqp2(torch_cuda, n, r):
    from oracle import clib
    Ur = _orth(n, r, 7 * n + r)
    o = clib.qrcp_dlaqp2(Ur)
    piv, rdiag, gap = _gpu_qrcp(torch_cuda, Ur, block=1)
    np.testing.assert_array_equal(piv, o["piv"])
    if r <= 100:        # register-resident dlarf pass: the oracle's fma sequence, bit for bit
        np.testing.assert_array_equal(rdiag, o["rdiag"])
    else:               # taller trailing blocks run the same algorithm on the tensor path
        np.testing.assert_allclose(rdiag, o["rdiag"], rtol=1e-13)
    np.testing.assert_allclose(gap, o["gap"], rtol=0, atol=1e-12)
    _, _, P = sla.qr(Ur.T, pivoting=True, mode="economic")
    np.testing.assert_array_equal(piv, P[:r])                   # and LAPACK's pivots


@pytest.mark.parametrize("n,r,block", [(50, 5, 2), (2001, 14, 4), (5000, 40, 8), (6000, 100, 8),
                                       (6000, 100, 3), (3001, 128, 6), (100000, 24, 5)])
def test_qrcp_blocked_matches_lapack_pivots(torch_cuda, n, r, block):
    Ur = _orth(n, r, 7 * n + r)
    _, R, P = sla.qr(Ur.T, pivoting=True, mode="economic")
    piv, rdiag, _ = _gpu_qrcp(torch_cuda, Ur, block=block)
    np.testing.assert_array_equal(piv, P[:r])
    np.testing.assert_allclose(np.abs(rdiag), np.abs(np.diag(R)), rtol=1e-11)


@pytest.mark.parametrize("block", [1, 4])
def test_qrcp_exact_ties_follow_lapack_order(torch_cuda, block):
    for seed in range(6):
        rng = np.random.default_rng(seed)
        base = _orth(60, 6, seed)
        Ur = np.concatenate([base, base[rng.permutation(60)[:30]], base], axis=0)
        _, _, P = sla.qr(Ur.T, pivoting=True, mode="economic")
        piv, _, gap = _gpu_qrcp(torch_cuda, Ur, block=block)
        np.testing.assert_array_equal(piv, P[:6])
        assert gap.min() == 0.0                                 # the meter reports the ties


def _localised_basis(n, r, seed):
    """Orthonormal n x r basis whose row norms vary by orders of magnitude along the mesh (localised
    modes, as flame data has them) -- the case in which most segments sit the in-block passes out."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 1.0, n)[:, None]
    centres, widths = rng.random(r)[None, :], 0.02 + 0.2 * rng.random(r)[None, :]
    A = np.exp(-((x - centres) / widths) ** 2) * np.cos(2 * np.pi * x * (1 + np.arange(r))[None, :] + rng.random(r))
    A += 1e-3 * rng.standard_normal((n, r))
    Q, _ = np.linalg.qr(A)
    return Q


@pytest.mark.parametrize("n,r,block,s", [(50, 5, 2, None), (63, 6, 8, None), (129, 8, 3, None), (2001, 14, 4, None),
                                         (5000, 40, 8, None), (6000, 100, 8, None), (6000, 100, 3, 37),
                                         (3001, 128, 6, None), (100000, 24, 5, None), (40000, 130, 8, 64),
                                         (250000, 40, 8, None)])
@pytest.mark.parametrize("kind", ["flat", "localised"])
def test_qrcp_lazy_downdates_match_eager_and_lapack(torch_cuda, n, r, block, s, kind):
    """Lazy norm down-dates (skipped segments, catch-up rounds, deferred down-dates in the apply pass)
    pick the pivots of the eager schedule = LAPACK's, for a loose, the default and an over-eager bound."""
    Ur = _orth(n, r, 3 * n + r) if kind == "flat" else _localised_basis(n, r, n + r)
    _, R, P = sla.qr(Ur.T, pivoting=True, mode="economic")
    k = r if s is None else s
    st0 = {}
    piv0, rd0, gap0 = _gpu_qrcp(torch_cuda, Ur, block=block, s=s, lazy=0.0, stats=st0)
    np.testing.assert_array_equal(piv0, P[:k])
    assert not st0["lazy"] and st0["retries"] == 0
    retries = 0
    for alpha in (0.5, 0.94, 0.9995):
        st = {}
        piv, rd, gap = _gpu_qrcp(torch_cuda, Ur, block=block, s=s, lazy=alpha, stats=st)
        np.testing.assert_array_equal(piv, piv0)
        np.testing.assert_allclose(rd, rd0, rtol=1e-12)
        np.testing.assert_allclose(np.abs(rd), np.abs(np.diag(R))[:k], rtol=1e-11)
        assert st["lazy"] and st["seg_rows"] <= st0["seg_rows"] and np.all(gap <= gap0 + 1e-12)
        retries += st["retries"]
        if kind == "localised" and n >= 5000 and alpha == 0.94:
            assert st["seg_rows"] < 0.9 * st0["seg_rows"]        # part of the mesh is never read inside a block
    if n >= 2001 and block > 2:
        assert retries > 0                                         # the catch-up path has run


def test_qrcp_apply_kernels_agree_bitwise(torch_cuda, tmp_path):
    """The block-closing apply pass has two implementations (tensor-copy / TMA landing stages for 41..104
    trailing rows, register-staged otherwise): same products, same norm arithmetic -> the same bits.  The
    switch is read once per process, so the register-staged run happens in a child process."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    Ur = _localised_basis(20000, 100, 5)
    np.save(tmp_path / "Ur.npy", Ur)
    code = ("import sys, numpy as np, torch; sys.path.insert(0, %r)\n"
            "from openmeasure_b200 import engine\n"
            "Ur = np.load(%r)\n"
            "eng = engine.Engine(torch.zeros(Ur.shape[0], 1, dtype=torch.float64, device='cuda'), 1, group=False)\n"
            "eng.set_basis_rows(torch.from_numpy(Ur).cuda())\n"
            "piv, rd, gap = eng.qrcp(block=8)\n"
            "np.savez(%r, piv=piv.cpu().numpy(), rd=rd.cpu().numpy(), gap=gap.cpu().numpy())\n"
            % (root, str(tmp_path / "Ur.npy"), str(tmp_path / "reg.npz")))
    env = dict(os.environ, OMB_QR_APPLY_TMA="0")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    reg = np.load(tmp_path / "reg.npz")
    piv, rd, gap = _gpu_qrcp(torch_cuda, Ur, block=8)
    np.testing.assert_array_equal(piv, reg["piv"])
    np.testing.assert_array_equal(rd, reg["rd"])
    np.testing.assert_array_equal(gap, reg["gap"])
    _, _, P = sla.qr(Ur.T, pivoting=True, mode="economic")
    np.testing.assert_array_equal(piv, P[:100])


def test_qrcp_lazy_ties_and_zero_columns(torch_cuda):
    """Exact ties across segments and masked (zero) candidates under an over-eager bound."""
    for seed in range(4):
        rng = np.random.default_rng(seed)
        base = _orth(300, 7, seed)
        Ur = np.concatenate([base, base[rng.permutation(300)[:170]], 0.0 * base[:90], base], axis=0)
        _, _, P = sla.qr(Ur.T, pivoting=True, mode="economic")
        for alpha in (0.94, 0.9995):
            piv, _, gap = _gpu_qrcp(torch_cuda, Ur, block=4, lazy=alpha)
            np.testing.assert_array_equal(piv, P[:7])
            assert gap.min() == 0.0


def test_qrcp_partial_steps_and_rank_deficient_columns(torch_cuda):
    from oracle import clib
    Ur = _orth(900, 12, 11)
    Ur[100:400] = 0.0                                           # masked-out candidates
    o = clib.qrcp_dlaqp2(Ur)
    piv, rdiag, _ = _gpu_qrcp(torch_cuda, Ur, block=1)
    np.testing.assert_array_equal(piv, o["piv"])
    piv5, _, _ = _gpu_qrcp(torch_cuda, Ur, block=3, s=5)
    np.testing.assert_array_equal(piv5, o["piv"][:5])


# ---------------------------------------------------------------------------------------------
# whole path vs the reference's golden fixtures (tests/golden, oracle/make_golden.py)
# ---------------------------------------------------------------------------------------------
def test_pipeline_matches_reference_golden(torch_cuda, golden):
    g = golden
    X = g["X"]
    n, m = X.shape
    n_c = n // g["F"]
    spr = _sps().SPR(X.copy(), g["F"], np.zeros((n_c, 3)))
    spr.fit(scale_type=g["scale_type"], axis_cnt=g["axis_cnt"], select_modes=g["select_modes"],
            n_modes=g["n_modes"])
    np.testing.assert_array_equal(spr.X_cnt, g["X_cnt"])
    np.testing.assert_array_equal(spr.X_scl, g["X_scl"])
    np.testing.assert_array_equal(spr.X0[:64], g["X0_head"])
    assert spr.r == g["r"]
    np.testing.assert_allclose(spr.Sigma_r, g["Sigma_r"], rtol=RTOL)
    Ur, sgn = _sign_align(spr.Ur, g["Ur"])
    # modes up to sign; a mode's accuracy scales with its neighbouring singular-value gap
    np.testing.assert_allclose(Ur, g["Ur"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(spr.Ar * sgn, g["Ar"], rtol=0, atol=1e-9 * g["Sigma_r"][0])
    C = spr.optimal_placement()
    assert C.shape == (g["r"], n)
    np.testing.assert_array_equal(spr.qr_pivots, g["piv"])      # bit-exact pivots
    assert spr.qr_gap.min() > 1e-9
    spr.train(C)
    np.testing.assert_allclose(spr.Theta * sgn, g["Theta"], rtol=0, atol=1e-9)
    Ar, Asig = spr.predict(list(g["Y"]))
    Xrec = spr.reconstruct(Ar)
    np.testing.assert_allclose(Xrec, g["X_rec"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(Ar * sgn, g["Ar_pred"], rtol=0, atol=1e-8 * np.abs(g["Ar_pred"]).max())
    np.testing.assert_allclose(Asig, g["Ar_sigma"], rtol=0, atol=1e-8 * max(np.abs(g["Ar_sigma"]).max(), 1e-300))


# ---------------------------------------------------------------------------------------------
# mirrors of the reference's own unit tests (tests/test_rom.py, tests/test_spr.py), same shapes
# ---------------------------------------------------------------------------------------------
class TestReferenceUnitTests:
    def setup_method(self, method):
        rng = np.random.default_rng(12345)
        self.n_points, self.n_features, self.m = 10, 2, 5
        self.X = rng.random(size=(self.n_points * self.n_features, self.m))
        self.xyz = rng.random(size=(self.n_points, 3))
        self.C = np.eye(self.X.shape[0])

    def test_decomposition_svd(self, torch_cuda):           # test_rom.py:48-55, up to sign
        rom = _sps().ROM(self.X, self.n_features, self.xyz)
        X0 = rom.scale_data()
        U, Sigma, Vt = np.linalg.svd(X0, full_matrices=False)
        A = np.dot(np.diag(Sigma), Vt).T
        Ur, Ar, _ = rom.decomposition(X0, select_modes='number', n_modes=self.m - 1)
        Ur, sgn = _sign_align(Ur, U[:, :self.m - 1])
        np.testing.assert_allclose(Ur, U[:, :self.m - 1], atol=1e-10)
        np.testing.assert_allclose(Ar * sgn, A[:, :self.m - 1], atol=1e-10)

    def test_reduction_number_and_variance(self, torch_cuda):   # test_rom.py:57-65
        rom = _sps().ROM(self.X, self.n_features, self.xyz)
        X0 = rom.scale_data()
        rom.decomposition(X0, select_modes='number', n_modes=self.m - 1)
        assert rom.r == self.m - 1
        rom.decomposition(X0, select_modes='variance', n_modes=100)
        assert rom.r == self.m

    def test_fit(self, torch_cuda):                           # test_rom.py:67-74
        rom = _sps().ROM(self.X, self.n_features, self.xyz)
        X0 = rom.scale_data()
        _, Sigma, Vt = np.linalg.svd(X0, full_matrices=False)
        rom.fit(n_modes=100)
        k = self.m - 1                                        # mode m is rounding noise (row-centred)
        V, _ = _sign_align(rom.Vr[:, :k], Vt.T[:, :k])
        np.testing.assert_allclose(V, Vt.T[:, :k], atol=1e-9)
        np.testing.assert_allclose(rom.Sigma_r[:k], Sigma[:k], rtol=RTOL)

    def test_unscaling(self, torch_cuda):                     # test_rom.py:76-80
        rom = _sps().ROM(self.X, self.n_features, self.xyz)
        X0 = rom.scale_data()
        rom.fit(n_modes=100)
        np.testing.assert_allclose(rom.unscale_data(X0[:, 0]), rom.X[:, 0])

    def test_reconstruction(self, torch_cuda):                # test_rom.py:82-85
        rom = _sps().ROM(self.X, self.n_features, self.xyz)
        rom.fit(n_modes=100)
        np.testing.assert_allclose(rom.reconstruct(rom.Ar[0, :]), rom.X[:, [0]], rtol=RTOL)

    def test_optimal_placement_qr(self, torch_cuda):          # test_spr.py:21-25
        spr = _sps().SPR(self.X, self.n_features, self.xyz)
        spr.fit(n_modes=100)
        C_qr = spr.optimal_placement()
        assert C_qr.shape[0] == self.m and C_qr.shape[1] == spr.X.shape[0]

    def test_scale_vector_and_predict(self, torch_cuda):      # test_spr.py:27-60
        spr = _sps().SPR(self.X, self.n_features, self.xyz)
        X_cnt = np.mean(self.X, axis=1)[:, np.newaxis]
        X_scl = np.zeros((self.X.shape[0], 1))
        for f in range(self.n_features):
            sl = slice(f * self.n_points, (f + 1) * self.n_points)
            X_scl[sl] = np.std(self.X[sl])
        spr.fit(n_modes=100)
        spr.train(self.C)
        y = np.zeros((self.C.shape[0], 3))
        y[:, 0] = self.C @ self.X[:, 0]
        for f in range(self.n_features):
            y[f * self.n_points:(f + 1) * self.n_points, 2] = f
        y0 = spr.scale_vector(y)
        y0_check = np.zeros((self.C.shape[0], 2))
        y0_check[:, 0] = (y[:, 0] - X_cnt[:, 0]) / X_scl[:, 0]
        np.testing.assert_allclose(y0, y0_check)
        a, _ = spr.predict(y)
        np.testing.assert_allclose(spr.reconstruct(a), self.X[:, [0]])

    def test_errors(self, torch_cuda):                        # sparse_sensing.py:752-754, :791-803, :848-854
        spr = _sps().SPR(self.X, self.n_features, self.xyz)
        spr.fit(select_modes='number', n_modes=3)
        with pytest.raises(NotImplementedError):
            spr.optimal_placement(calc_type='bogus')
        with pytest.raises(ValueError):
            spr.train(np.eye(7))
        spr.train(spr.optimal_placement())
        with pytest.raises(ValueError):
            spr.predict(np.zeros((4, 3)))
        with pytest.raises(ValueError):
            spr.predict(np.zeros((3, 2)))


# ---------------------------------------------------------------------------------------------
# mask quirk (the basis is zeroed in place), general dense C, basis= resume hook
# ---------------------------------------------------------------------------------------------
def test_mask_basis_dense_c(torch_cuda):
    from oracle import pod_oracle as po, synth as osynth
    X = osynth.snapshots(3, 500, 20, 8)
    n = X.shape[0]
    f = po.fit(X, 3, n_modes=8, select_modes="number")
    spr = _sps().SPR(X, 3, None)
    spr.fit(select_modes='number', n_modes=8)
    mask = np.ones(n, dtype=bool)
    mask[200:900] = False
    piv_ref = po.qr_pivots(f["Ur"], mask)
    C = spr.optimal_placement(mask=mask)
    np.testing.assert_array_equal(np.argmax(np.asarray(C), axis=1), piv_ref)
    assert np.all(spr.Ur[~mask] == 0)                          # reference quirk :737-738
    # general (non one-hot) dense C goes through the dense product
    rng = np.random.default_rng(0)
    Cd = rng.random((12, n))
    spr.train(Cd)
    np.testing.assert_allclose(spr.Theta, Cd @ spr.Ur, rtol=1e-12, atol=1e-14)
    # resume from a given basis (sparse_sensing.py:493-497)
    spr2 = _sps().SPR(X, 3, None)
    spr2.fit(basis=(f["Ur"], f["Ar"]))
    assert spr2.r == 8
    np.testing.assert_array_equal(spr2.Ur, f["Ur"])
    C2 = spr2.optimal_placement(block=1)
    np.testing.assert_array_equal(spr2.qr_pivots, po.qr_pivots(f["Ur"]))


# ---------------------------------------------------------------------------------------------
# size-independent properties at a config-2-like size (oracle too slow to run in full here)
# ---------------------------------------------------------------------------------------------
def test_large_properties(torch_cuda):
    from openmeasure_b200 import synth as gsynth
    from oracle import clib
    torch = torch_cuda
    F, n_c, m, r = 9, 60000, 41, 40
    Xd = gsynth.snapshots(F, n_c, m, r)
    spr = _sps().SPR.from_device(Xd, F)
    spr.fit(select_modes='number', n_modes=r)
    U = spr._eng.basis_rows()
    G = (U.T @ U).cpu().numpy()
    np.testing.assert_allclose(G, np.eye(r), atol=1e-9)        # orthonormal modes
    assert spr.pod_rel_err_bound < 1e-10
    C = spr.optimal_placement()
    o = clib.qrcp_dlaqp2(U.cpu().numpy())                       # oracle QRCP on the GPU's own basis
    np.testing.assert_array_equal(spr.qr_pivots, o["piv"])
    np.testing.assert_allclose(np.abs(spr.qr_rdiag), np.abs(o["rdiag"]), rtol=1e-11)
    spr.train(C)
    # snapshot 3 sampled at the sensors -> reconstruct: error bounded by the truncated energy
    x = Xd[:, 3].cpu().numpy()
    y = np.zeros((r, 3))
    y[:, 0] = x[spr.qr_pivots]
    y[:, 2] = spr.qr_pivots // n_c
    a, _ = spr.predict(y)
    xr = spr.reconstruct(a)[:, 0]
    assert np.linalg.norm(xr - x) / np.linalg.norm(x) < 1e-3
    # linearity of predict/reconstruct in the measurements (about the centring)
    y2 = y.copy()
    y2[:, 0] = 2 * y[:, 0] - spr.X_cnt[spr.qr_pivots, 0]
    a2, _ = spr.predict(y2)
    np.testing.assert_allclose(a2, 2 * a, rtol=1e-9, atol=1e-9 * np.abs(a).max())


# ---------------------------------------------------------------------------------------------
# row-sharded (multi-rank) path: G ranks emulated by G threads on this one GPU (ThreadComm), every
# kernel issued on the same stream.  Results must not depend on the number of ranks.
# ---------------------------------------------------------------------------------------------
def _run_ranks(torch, X, F, cells, r, block):
    import threading
    from openmeasure_b200 import comm as C
    n_c = sum(cells)
    comms = C.ThreadComm.make(len(cells))
    out = [None] * len(cells)
    err = []

    def run(rk):
        try:
            lay = C.ShardLayout(F, cells, rk)
            rows = lay.to_global(torch.arange(F * cells[rk])).numpy()
            Xl = torch.from_numpy(np.ascontiguousarray(X[rows])).cuda()
            spr = _sps().SPR.from_device(Xl, F, comm=comms[rk])
            spr.fit(select_modes='number', n_modes=r)
            Cm = spr.optimal_placement(block=block)
            spr.train(Cm)
            y = np.zeros((r, 3))
            y[:, 0] = X[Cm.pivots, 2]
            y[:, 2] = Cm.pivots // n_c
            a, _ = spr.predict(y)
            out[rk] = dict(piv=Cm.pivots.copy(), S=spr.Sigma_r.copy(), Theta=spr.Theta.copy(), a=a,
                           rec=spr.reconstruct(a), rows=rows, shape=Cm.shape, scl=spr.X_scl[::cells[rk], 0].copy())
        except Exception as e:          # pragma: no cover
            err.append(e)
            comms[rk].shared.barrier.abort()

    th = [threading.Thread(target=run, args=(k,)) for k in range(len(cells))]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    return out


@pytest.mark.parametrize("cells,block", [([400, 300], 4), ([200, 260, 240], 1), ([350, 350], 8)])
def test_multirank_matches_single_rank(torch_cuda, cells, block):
    from oracle import pod_oracle as po, synth as osynth
    F, m, r = 3, 24, 10
    n_c = sum(cells)
    X = osynth.snapshots(F, n_c, m, r)
    ref = po.placement_pipeline(X, F, r)
    one = _run_ranks(torch_cuda, X, F, [n_c], r, block)[0]
    np.testing.assert_array_equal(one["piv"], ref["piv"])
    outs = _run_ranks(torch_cuda, X, F, cells, r, block)
    full = np.zeros((F * n_c, 1))
    for o in outs:
        assert o["shape"] == (r, F * n_c)
        np.testing.assert_array_equal(o["piv"], ref["piv"])                 # global pivots, every rank
        np.testing.assert_allclose(o["S"], ref["Sigma_r"], rtol=RTOL)
        np.testing.assert_allclose(o["scl"], ref["X_scl"][::n_c, 0], rtol=1e-14)
        np.testing.assert_allclose(np.abs(o["Theta"]), np.abs(one["Theta"]), rtol=0, atol=1e-10)
        full[o["rows"]] = o["rec"]
    np.testing.assert_allclose(full, one["rec"], rtol=1e-9)


@pytest.mark.parametrize("cells,m,r", [([300, 260], 130, 40), ([200, 180, 190], 256, 64), ([333, 300], 256, 100)])
def test_multirank_many_snapshots(torch_cuda, cells, m, r):
    """m > 64: the many-snapshot Gram / back-projection kernels and the library eigensolver whose result rank 0
    broadcasts (engine.eig_pod) -- the branch configs[2] and configs[4] take on more than one GPU."""
    from oracle import pod_oracle as po, synth as osynth
    F = 2
    n_c = sum(cells)
    X = osynth.snapshots(F, n_c, m, r)
    ref = po.placement_pipeline(X, F, r)
    one = _run_ranks(torch_cuda, X, F, [n_c], r, 8)[0]
    np.testing.assert_array_equal(one["piv"], ref["piv"])
    np.testing.assert_allclose(one["S"], ref["Sigma_r"], rtol=RTOL)
    full = np.zeros((F * n_c, 1))
    for o in _run_ranks(torch_cuda, X, F, cells, r, 8):
        np.testing.assert_array_equal(o["piv"], ref["piv"])
        np.testing.assert_allclose(o["S"], one["S"], rtol=1e-12)             # G-invariance of sigma
        np.testing.assert_allclose(np.abs(o["Theta"]), np.abs(one["Theta"]), rtol=0, atol=1e-10)
        full[o["rows"]] = o["rec"]
    np.testing.assert_allclose(full, one["rec"], rtol=1e-9)


def test_two_gpu_peer_memory_parity(torch_cuda):
    """Real NVLink peers (needs >= 2 GPUs; the driver's 1-GPU run skips it): torchrun tools/mr_check.py -- the
    peer-memory all-gather / all-reduce kernels and the in-kernel pivot exchange against the single-GPU run."""
    import os, subprocess, sys
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # OMB_QR_LAZY: the default bound, an over-eager one (a catch-up round and a second exchange at almost
    # every step) and the eager schedule
    for extra, lazy in ((["20000", "41", "40"], None), (["6000", "256", "100"], None), (["20000", "41", "40"], "0.9995"),
                        (["6000", "256", "100"], "0.9995"), (["20000", "41", "40"], "0")):
        env = dict(os.environ)
        if lazy is not None:
            env["OMB_QR_LAZY"] = lazy
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                              "--master-addr", "127.0.0.1", "--master-port", "29533",
                              os.path.join(root, "tools", "mr_check.py")] + extra,
                             capture_output=True, text=True, timeout=600, env=env)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "-> OK" in out.stdout
        if lazy == "0.9995":
            assert "catch_up_rounds=0 " not in out.stdout


# ---------------------------------------------------------------------------------------------
# S3: one-CTA Jacobi eigensolver vs LAPACK (np.linalg.eigh), including the rank-deficient Gram of
# row-centred data (sigma_m ~ 0) and graded spectra
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [1, 2, 5, 14, 41, 64])
@pytest.mark.parametrize("kind", ["gram", "graded"])
def test_eigh_jacobi_matches_lapack(torch_cuda, m, kind):
    import ctypes as C
    from openmeasure_b200 import _lib
    torch = torch_cuda
    rng = np.random.default_rng(100 * m + len(kind))
    if kind == "gram":
        X = rng.standard_normal((500, m))
        X -= X.mean(axis=1, keepdims=True)          # row-centred: rank m - 1
        G = X.T @ X
    else:
        Q, _ = np.linalg.qr(rng.standard_normal((m, m)))
        G = (Q * np.logspace(0, -10, m)) @ Q.T
        G = (G + G.T) / 2
    Gd = torch.from_numpy(G).cuda()
    w = torch.empty(m, dtype=torch.float64, device="cuda")
    V = torch.empty(m, m, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("omb_eigh_jacobi", C.c_void_p(Gd.data_ptr()), m, C.c_void_p(w.data_ptr()), C.c_void_p(V.data_ptr()),
              C.c_void_p(info.data_ptr()), None)
    w, V = w.cpu().numpy(), V.cpu().numpy()
    wl = np.linalg.eigvalsh(G)[::-1]
    scale = np.abs(wl).max()
    np.testing.assert_allclose(w, wl, rtol=0, atol=4e-15 * scale * max(m, 4))
    np.testing.assert_allclose(V.T @ V, np.eye(m), atol=1e-13)
    np.testing.assert_allclose(G @ V, V * w, atol=1e-13 * scale * max(m, 4))
    assert np.all(np.diff(w) <= 0) and 0 < int(info.item()) + 1 <= 30


# ---------------------------------------------------------------------------------------------
# K3/K5: Gram and back-projection kernels against numpy over the kernel-selection boundaries
# (m <= 64 register-resident variant with fused row means; even m > 64 TMA-pipelined 128 x 128
# tiles; odd m > 64 staged 64 x 64 tiles) and ragged row counts
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,n_c,m,r", [(2, 37, 8, 3), (3, 1000, 41, 14), (2, 515, 64, 20), (2, 700, 66, 30),
                                       (2, 333, 128, 64), (1, 2049, 130, 100), (3, 900, 256, 100),
                                       (2, 450, 300, 128), (2, 611, 129, 40), (1, 300, 640, 200)])
def test_gram_and_backprojection_match_numpy(torch_cuda, F, n_c, m, r):
    torch = torch_cuda
    from openmeasure_b200 import engine as E
    rng = np.random.default_rng(F * 1000 + m)
    X = rng.random((F * n_c, m)) * 10.0 ** rng.integers(-1, 2, (F * n_c, 1)) + 3.0
    eng = E.Engine(torch.from_numpy(X).cuda(), F, group=False)
    eng.stats("std", 1, defer_row_means=True)
    G = eng.gram().cpu().numpy()
    np.testing.assert_array_equal(eng.cnt.cpu().numpy(), np.mean(X, axis=1))       # fused row means: bit-exact
    scl = np.repeat([np.std(X[f * n_c:(f + 1) * n_c]) for f in range(F)], n_c)[:, None]
    X0 = (X - np.mean(X, axis=1)[:, None]) / scl
    Gref = X0.T @ X0
    np.testing.assert_allclose(G, Gref, rtol=0, atol=2e-13 * np.abs(Gref).max())
    np.testing.assert_array_equal(G, G.T)
    W = rng.standard_normal((m, r))
    eng.backproject(torch.from_numpy(W).cuda())
    U = eng.basis_rows().cpu().numpy()
    Uref = X0 @ W
    np.testing.assert_allclose(U, Uref, rtol=0, atol=2e-13 * np.abs(Uref).max())
    np.testing.assert_allclose(eng.vn.cpu().numpy()[:F * n_c], np.linalg.norm(Uref, axis=1), rtol=1e-12)


@pytest.mark.parametrize("seed", range(36))
def test_many_snapshot_kernels_random_shapes(torch_cuda, seed):
    """The stream-K row slices of the Gram (a CTA's slice may start and end anywhere: in the middle of a tile's rows,
    across tiles and features, or be empty), ragged last chunks, ragged last column tiles and single-chunk blocks;
    the centred copy; the 16-row-warp back-projection with every mode-block count -- against numpy."""
    torch = torch_cuda
    from openmeasure_b200 import engine as E
    rng = np.random.default_rng(1000 + seed)
    F = int(rng.integers(1, 5))
    m = int(rng.choice([66, 70, 128, 130, 200, 256, 258, 384, 520, 67, 129, 257, 301]))     # odd counts: padded copy
    n_c = int(rng.choice([1, 3, 15, 16, 17, 40, 129, 500, 1237, 2900]))
    r = int(rng.integers(1, min(m, 140) + 1)) if seed % 3 == 0 else int(2 * rng.integers(1, min(m, 140) // 2 + 1))
    X = rng.standard_normal((F * n_c, m)) * 10.0 ** rng.integers(-2, 3, (F * n_c, 1)) + rng.standard_normal((F * n_c, 1))
    if n_c * m < 2:
        pytest.skip("degenerate block")
    eng = E.Engine(torch.from_numpy(X).cuda(), F, group=False)
    eng.stats("none", 1, defer_row_means=True)
    G = eng.gram().cpu().numpy()
    np.testing.assert_array_equal(eng.cnt.cpu().numpy(), np.mean(X, axis=1))
    X0 = X - np.mean(X, axis=1)[:, None]
    Gref = X0.T @ X0
    np.testing.assert_allclose(G, Gref, rtol=0, atol=5e-13 * max(np.abs(Gref).max(), 1e-300))
    np.testing.assert_array_equal(G, G.T)
    W = rng.standard_normal((m, r))
    eng.backproject(torch.from_numpy(W).cuda())
    Uref = X0 @ W
    np.testing.assert_allclose(eng.basis_rows().cpu().numpy(), Uref, rtol=0, atol=5e-13 * max(np.abs(Uref).max(), 1e-300))
    np.testing.assert_allclose(eng.vn.cpu().numpy()[:F * n_c], np.linalg.norm(Uref, axis=1), rtol=1e-11, atol=1e-300)
    # the same without room for the centred copy: in-kernel centring (centring warps / CENTRE = true instantiations)
    import os
    os.environ["OMB_CENTRED_COPY"] = "0"
    try:
        eng2 = E.Engine(torch.from_numpy(X).cuda(), F, group=False)
        eng2.stats("none", 1, defer_row_means=True)
        G2 = eng2.gram().cpu().numpy()
        eng2.backproject(torch.from_numpy(W).cuda())
        U2 = eng2.basis_rows().cpu().numpy()
    finally:
        del os.environ["OMB_CENTRED_COPY"]
    np.testing.assert_allclose(G2, Gref, rtol=0, atol=5e-13 * max(np.abs(Gref).max(), 1e-300))
    np.testing.assert_allclose(U2, Uref, rtol=0, atol=5e-13 * max(np.abs(Uref).max(), 1e-300))


# ---------------------------------------------------------------------------------------------
# GEM placement: the reference's own selections (golden g6/g7, jitter reproduced
# by seeding numpy like the fixture generator) and the oracle on a larger case
# ---------------------------------------------------------------------------------------------
def test_gem_matches_reference_golden(torch_cuda, golden_gem):
    g = golden_gem
    F, n_c = int(g["F"]), g["X"].shape[0] // int(g["F"])
    spr = _sps().SPR(g["X"], F, g["xyz"])
    spr.fit(select_modes="number", n_modes=int(g["n_modes"]))
    # GEM takes variances ACROSS the modes of a row, so -- unlike the QR placement -- its result depends
    # on the arbitrary signs of the singular vectors: parity is defined on the reference's own basis
    s_ref, s_own = np.linalg.svd(g["Ur"].T @ spr.Ur, compute_uv=False), None
    np.testing.assert_allclose(s_ref, 1.0, atol=1e-8)                    # same subspace
    spr.Ur = g["Ur"]
    np.random.seed(int(g["seed"]))
    C = spr.optimal_placement(calc_type="gem", n_sensors=int(g["n_sensors"]), mask=g["mask"], d_min=float(g["d_min"]))
    assert C.shape == (int(g["n_sensors"]), F * n_c)
    np.testing.assert_array_equal(C.pivots, g["gem"])
    # explicit basis argument (the reference signature gem(Ur, n_sensors, mask, d_min, verbose))
    np.random.seed(int(g["seed"]))
    P = spr.gem(g["Ur"], int(g["n_sensors"]), g["mask"], float(g["d_min"]), False)
    np.testing.assert_array_equal(P, g["gem"])


@pytest.mark.parametrize("n_sensors,d_min", [(12, 0.0), (20, 0.05), (40, 0.0)])
def test_gem_matches_oracle_larger(torch_cuda, n_sensors, d_min):
    from oracle import pod_oracle as po, synth as osynth
    F, n_c, m, r = 3, 7000, 48, 24
    X = osynth.snapshots(F, n_c, m, r)
    rng = np.random.default_rng(3)
    xyz = rng.random((n_c, 3))
    mask = rng.random(F * n_c) > 0.2
    spr = _sps().SPR(X, F, xyz)
    spr.fit(select_modes="number", n_modes=r)
    draws = [rng.standard_normal(k) for k in range(n_sensors + 1)]      # the same jitter for both sides
    take = lambda seq: (lambda size: seq.pop(0) if len(seq[0]) == size else seq.pop(0)[:size])
    seq1 = [d.copy() for d in draws[2:]]
    seq2 = [d.copy() for d in draws[2:]]
    ref, cond = po.gem_placement(spr.Ur, xyz, F, n_sensors, mask, d_min, normal=take(seq1))
    eng = spr._eng
    import torch
    got = eng.gem(n_sensors, torch.from_numpy(mask).cuda(), torch.from_numpy(xyz).cuda(), d_min, normal=take(seq2))
    np.testing.assert_array_equal(got, ref)
    assert np.all(mask[got])
    if n_sensors < r:       # beyond r - 1 chosen rows the covariance is singular (jitter only): repeats happen, as in the reference
        assert len(set(got.tolist())) == n_sensors


# ---------------------------------------------------------------------------------------------
# reconstruct(..., sampling=S) / unscale_data(x0, sampling=S) (sparse_sensing.py:233, :365-368):
# one-hot, dense and scipy-sparse sampling matrices against the reference formulas in numpy
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["onehot", "sensor_matrix", "dense", "sparse"])
def test_reconstruct_with_sampling(torch_cuda, kind):
    import scipy.sparse as sp
    from oracle import synth as osynth
    sps = _sps()
    F, n_c, m, r = 3, 500, 20, 8
    X = osynth.snapshots(F, n_c, m, r)
    n = F * n_c
    spr = sps.SPR(X, F, np.zeros((n_c, 3)))
    spr.fit(select_modes="number", n_modes=r)
    rng = np.random.default_rng(9)
    rows = rng.choice(n, 17, replace=False)
    S = np.zeros((17, n))
    S[np.arange(17), rows] = 1
    if kind == "sensor_matrix":
        Sarg = sps.SensorMatrix(rows, n)
    elif kind == "dense":
        S = rng.random((17, n)) * (rng.random((17, n)) < 0.01)
        Sarg = S
    elif kind == "sparse":
        S = rng.random((17, n)) * (rng.random((17, n)) < 0.01)
        Sarg = sp.csr_matrix(S)
    else:
        Sarg = S
    Ar = rng.standard_normal((5, r))
    ref = (S @ spr.X_scl[:, 0])[:, None] * np.linalg.multi_dot([S, spr.Ur, Ar.T]) + (S @ spr.X_cnt[:, 0])[:, None]
    got = spr.reconstruct(Ar, sampling=Sarg)
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-12 * np.abs(ref).max())
    x0 = rng.standard_normal(17)
    np.testing.assert_allclose(spr.unscale_data(x0, sampling=Sarg), (S @ spr.X_scl[:, 0]) * x0 + S @ spr.X_cnt[:, 0],
                               rtol=1e-13)
    if kind == "onehot":                                  # consistent with sampling the full reconstruction
        np.testing.assert_allclose(got, spr.reconstruct(Ar)[rows], rtol=RTOL)


# ---------------------------------------------------------------------------------------------
# ingest: .npy snapshot file -> HBM shard (pinned ring + copy stream), then the same pipeline
# ---------------------------------------------------------------------------------------------
def test_npy_ingest_roundtrip_and_fit(torch_cuda, tmp_path):
    from openmeasure_b200 import ingest
    from oracle import pod_oracle as po, synth as osynth
    F, n_c, m, r = 3, 4001, 24, 10
    X = osynth.snapshots(F, n_c, m, r)
    path = str(tmp_path / "X_train.npy")
    np.save(path, X)
    for world in (1, 3):
        parts = [ingest.load_npy_shard(path, F, rk, world, chunk_bytes=100_000).cpu().numpy() for rk in range(world)]
        for rk, P in enumerate(parts):
            c0, ncl = ingest.shard_cells(n_c, rk, world)
            np.testing.assert_array_equal(P, np.concatenate([X[f * n_c + c0: f * n_c + c0 + ncl] for f in range(F)]))
    spr = _sps().SPR.from_npy(path, F, np.zeros((n_c, 3)), group=False)
    spr.fit(select_modes="number", n_modes=r)
    spr.optimal_placement()
    ref = po.placement_pipeline(X, F, r)
    np.testing.assert_array_equal(spr.X_cnt, ref["X_cnt"])
    np.testing.assert_array_equal(spr.qr_pivots, ref["piv"])


# ---------------------------------------------------------------------------------------------
# BASELINE.json's full sizes (configs[1]: 1.65M x 41, r = 40; configs[2]: 16.2M x 256, r = 100; configs[4]'s
# snapshot count, 1024, on a quarter of one GPU's shard): size-independent properties, everything checked on the device
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,n_c,m,r", [(9, 183620, 41, 40), (9, 1800000, 256, 100), (8, 262144, 1024, 100)])
def test_full_size_properties(torch_cuda, F, n_c, m, r):
    from openmeasure_b200 import synth as gsynth
    torch = torch_cuda
    free, _ = torch.cuda.mem_get_info()
    if free < 12 * F * n_c * m * 8 // 4:
        pytest.skip("not enough free HBM for this configuration")
    Xd = gsynth.snapshots(F, n_c, m, r)
    spr = _sps().SPR.from_device(Xd, F)
    spr.fit(select_modes='number', n_modes=r)
    eng = spr._eng
    n = F * n_c
    U = eng.basis_rows()
    # (1) orthonormal modes, (2) the POD identity  X0^T U = V Sigma  column by column (X0 never materialised:
    #     X0^T U = (X^T (U / scl_row) - ones * cnt^T (U / scl_row)))
    G = U.T @ U
    assert float((G - torch.eye(r, dtype=torch.float64, device=G.device)).abs().max()) < 1e-9
    scl_rows = torch.repeat_interleave(eng.scl, eng.n_c_loc)
    Us = U / scl_rows[:, None]
    B = Xd.T @ Us - torch.outer(torch.ones(m, dtype=torch.float64, device=U.device), eng.cnt @ Us)
    S = torch.from_numpy(spr.Sigma_r).to(U.device)
    Vr = torch.from_numpy(spr.Vr).to(U.device)
    assert float((B - Vr * S).abs().max() / S[0]) < 1e-9
    assert bool((S[:-1] >= S[1:]).all()) and spr.pod_rel_err_bound < 1e-10
    del Us, B, G
    # (3) placement: the blocked schedule picks the pivots of the unblocked (dlaqp2-bitwise) one, all distinct
    C = spr.optimal_placement(block=8)
    piv8 = spr.qr_pivots.copy()
    gap8 = float(spr.qr_gap.min())
    spr.optimal_placement(block=1)
    np.testing.assert_array_equal(piv8, spr.qr_pivots)
    assert len(set(piv8.tolist())) == r and gap8 > 1e-9 and piv8.min() >= 0 and piv8.max() < n
    # (4) sample -> predict -> reconstruct round trip on a field that lies in span(Ur): exact recovery
    spr.train(C)
    a_true = torch.linspace(-1.0, 1.0, r, dtype=torch.float64, device=U.device)
    x = scl_rows * (U @ a_true) + eng.cnt
    pv = torch.from_numpy(piv8).to(U.device)
    y = np.zeros((r, 3))
    y[:, 0] = x[pv].cpu().numpy()
    y[:, 2] = piv8 // n_c
    a, _ = spr.predict(y)
    np.testing.assert_allclose(a[0], a_true.cpu().numpy(), rtol=0, atol=1e-7)
    del U
    xr = eng.reconstruct(torch.from_numpy(a).to(Xd.device))[:, 0]
    assert float((xr - x).abs().max() / x.abs().max()) < 1e-8


# ---------------------------------------------------------------------------------------------
# POD accuracy on a graded spectrum (sigma_r / sigma_1 = 1e-5): the Gram route alone resolves the
# smallest singular values to ~1e-6; the CholeskyQR2-style basis correction restores 1e-10
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,m,r", [(6000, 16, 12), (4000, 96, 40)])
def test_pod_refinement_graded_spectrum(torch_cuda, n, m, r):
    sps = _sps()
    rng = np.random.default_rng(n + m)
    Uo, _ = np.linalg.qr(rng.standard_normal((n, m)))
    Vo, _ = np.linalg.qr(rng.standard_normal((m, m)))
    sig = np.concatenate([np.logspace(0, -5, r), np.logspace(-5.5, -7, m - r)])
    X0 = (Uo * sig) @ Vo.T
    Uref, Sref, Vtref = np.linalg.svd(X0, full_matrices=False)
    rom = sps.ROM(X0, 1, None)
    rom.pod_refine = False
    _, _, _ = rom.decomposition(X0, 'number', r)
    err_gram = np.max(np.abs(rom.pod_sigma[:r] - Sref[:r]) / Sref[:r])
    rom.pod_refine = 'auto'
    Ur, Ar, ev = rom.decomposition(X0, 'number', r)
    assert rom.pod_refined and rom.pod_rel_err_bound < 1e-10
    S = np.linalg.norm(Ar, axis=0)
    err_ref = np.max(np.abs(S - Sref[:r]) / Sref[:r])
    assert err_ref < 1e-10 < err_gram, (err_ref, err_gram)
    np.testing.assert_allclose(Ur.T @ Ur, np.eye(r), atol=1e-12)
    Ua, s = _sign_align(Ur, Uref[:, :r])
    # modes up to sign; the smallest ones are determined to eps * sigma_1 / (gap to the neighbours)
    np.testing.assert_allclose(Ua[:, : r // 2], Uref[:, : r // 2], atol=1e-9)
    np.testing.assert_allclose(Ur @ Ar.T, (Uref[:, :r] * Sref[:r]) @ Vtref[:r], atol=1e-10 * Sref[0])


@pytest.mark.parametrize("n,m,r,lo", [(6000, 16, 12, -6.0), (5000, 40, 30, -7.0), (4000, 96, 40, -7.0), (3000, 130, 60, -9.0)])
def test_pod_hard_spectrum_full_width_route(torch_cuda, n, m, r, lo):
    """sigma_r / sigma_1 = 1e-6 ... 1e-9 (SURVEY 8d's "hard" distribution): the Gram of X0 no longer holds the small
    modes (lambda_r/lambda_1 <= eps), the one-pass correction is not enough either; the full-width route must agree
    with the constructed spectrum and with np.linalg.svd (the reference's call, sparse_sensing.py:272) to 1e-10
    relative or to the eps * sigma_1 absolute accuracy ANY backward-stable SVD has -- and say so without a warning.
    With the correction switched off the same input must raise the accuracy warning."""
    import warnings
    sps = _sps()
    rng = np.random.default_rng(n + m)
    Uo, _ = np.linalg.qr(rng.standard_normal((n, m)))
    Vo, _ = np.linalg.qr(rng.standard_normal((m, m)))
    sig = np.concatenate([np.logspace(0, lo, r), np.logspace(lo - 0.5, lo - 2, m - r)])
    X0 = (Uo * sig) @ Vo.T
    Uref, Sref, Vtref = np.linalg.svd(X0, full_matrices=False)
    tol = np.maximum(1e-10 * sig[:r], 200 * np.finfo(float).eps * sig[0])
    rom = sps.ROM(X0, 1, None)
    with warnings.catch_warnings():
        warnings.filterwarnings("error", message="POD: estimated")
        Ur, Ar, ev = rom.decomposition(X0, 'number', r)
    assert rom.pod_refined == 'full' and rom.pod_rel_err_bound < 1e-10
    S = np.linalg.norm(Ar, axis=0)
    assert np.all(np.abs(S - sig[:r]) <= tol), np.max(np.abs(S - sig[:r]) / sig[:r])
    assert np.all(np.abs(S - Sref[:r]) <= tol)
    np.testing.assert_allclose(Ur.T @ Ur, np.eye(r), atol=1e-11)
    # the factorisation itself: U diag(sigma) V^T reproduces the retained part of X0 to eps * sigma_1
    np.testing.assert_allclose(Ur @ Ar.T, (Uref[:, :r] * Sref[:r]) @ Vtref[:r], atol=2e-13 * sig[0])
    Ua, _ = _sign_align(Ur, Uref[:, :r])
    np.testing.assert_allclose(Ua[:, : r // 3], Uref[:, : r // 3], atol=1e-9)
    rom.pod_refine = False
    with pytest.warns(RuntimeWarning, match="estimated relative error"):
        rom.decomposition(X0, 'number', r)


def test_pod_null_mode_of_centred_data_is_dropped_not_amplified(torch_cuda):
    """r = m on row-centred data: the structural null mode (sigma_m = 0 up to rounding) gets a zero weight -- a
    zero basis column, not rounding noise divided by ~1e-8 -- and does not trigger the full-width route."""
    import warnings
    from oracle import synth as osynth
    F, n_c, m = 3, 1500, 24
    X = osynth.snapshots(F, n_c, m, 16)
    spr = _sps().SPR(X, F, np.zeros((n_c, 3)))
    with warnings.catch_warnings():
        warnings.filterwarnings("error", message="POD: estimated")
        spr.fit(select_modes='number', n_modes=m)
    assert spr.pod_refined != 'full'
    U = spr.Ur
    assert np.abs(U[:, m - 1]).max() < 1e-6 and spr.Sigma_r[-1] < 1e-7 * spr.Sigma_r[0]
    np.testing.assert_allclose(U[:, :m - 1].T @ U[:, :m - 1], np.eye(m - 1), atol=1e-9)


# ---------------------------------------------------------------------------------------------
# batched reconstruct across the kernel-selection boundary (TMA-pipelined 128 x 128 tiles for even
# r / N, staged 64 x 64 tiles otherwise), ragged row and vector counts
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_c,r,N", [(700, 10, 7), (1000, 16, 16), (3001, 40, 130), (2050, 100, 64), (900, 14, 258)])
def test_batched_reconstruct_matches_numpy(torch_cuda, n_c, r, N):
    torch = torch_cuda
    from openmeasure_b200 import engine as E
    F, m = 2, max(r + 4, 24)
    rng = np.random.default_rng(n_c + r + N)
    X = rng.random((F * n_c, m)) + 1.0
    eng = E.Engine(torch.from_numpy(X).cuda(), F, group=False)
    eng.stats("std", 1)
    Ur = rng.standard_normal((F * n_c, r))
    eng.set_basis_rows(torch.from_numpy(Ur).cuda())
    A = rng.standard_normal((N, r))
    got = eng.reconstruct(torch.from_numpy(A).cuda()).cpu().numpy()
    scl = np.repeat(eng.scl.cpu().numpy(), n_c)[:, None]
    ref = scl * (Ur @ A.T) + eng.cnt.cpu().numpy()[:, None]
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 * np.abs(ref).max())


# ---------------------------------------------------------------------------------------------
# every scaling type x both centring modes against the reference's own outputs (fixture s1)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scale_type", ["std", "none", "pareto", "vast", "range", "level", "max", "variance",
                                        "median", "poisson", "l2-norm"])
@pytest.mark.parametrize("axis_cnt,tag", [(1, "row"), (None, "blk")])
def test_all_scalings_match_reference_golden(torch_cuda, scale_type, axis_cnt, tag):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "s1_scalings_450x11.npz"))
    rom = _sps().ROM(z["X"], int(z["F"]), None)
    X0 = rom.scale_data(scale_type, axis_cnt)
    key = scale_type.replace("-", "_") + "_" + tag
    np.testing.assert_array_equal(rom.X_cnt, z[key + "_cnt"])
    if scale_type in ("vast", "l2-norm"):
        # derived from two tree sums (sd^2 / mean; sqrt(q + N mean^2)): within 2 ulp of numpy's own route
        np.testing.assert_allclose(rom.X_scl, z[key + "_scl"], rtol=4e-16)
        np.testing.assert_allclose(X0, z[key + "_X0"], rtol=1e-14, atol=1e-14)
    else:
        np.testing.assert_array_equal(rom.X_scl, z[key + "_scl"])
        np.testing.assert_array_equal(X0, z[key + "_X0"])


# ---------------------------------------------------------------------------------------------
# API corners against the reference's own outputs (fixture x1): fit(basis=...), train with a general
# dense C and cond=True, weighted predict, reconstruct / unscale_data with a sampling matrix
# ---------------------------------------------------------------------------------------------
def test_api_extras_match_reference_golden(torch_cuda):
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "x1_extras_900x20_r8.npz"))
    F, r = int(z["F"]), int(z["r"])
    n_c = z["X"].shape[0] // F
    spr = _sps().SPR(z["X"], F, np.zeros((n_c, 3)))
    spr.fit(select_modes="number", n_modes=r, basis=(z["Ur"], z["Ar"]))     # the reference's own basis (signs)
    spr.train(z["C"], cond=True)
    np.testing.assert_allclose(spr.Theta, z["Theta"], rtol=1e-12, atol=1e-14)
    assert abs(spr.k - float(z["k"])) / float(z["k"]) < 1e-10
    Ar_p, Ar_sig = spr.predict(list(z["Y"]))
    np.testing.assert_allclose(Ar_p, z["Ar_pred"], rtol=1e-9, atol=1e-10 * np.abs(z["Ar_pred"]).max())
    np.testing.assert_allclose(Ar_sig, z["Ar_sigma"], rtol=1e-9, atol=1e-10 * np.abs(z["Ar_sigma"]).max())
    np.testing.assert_allclose(spr.cnt_vector, z["cnt_vector"], rtol=1e-13)
    np.testing.assert_array_equal(spr.scl_vector, z["scl_vector"])
    np.testing.assert_allclose(spr.reconstruct(Ar_p, sampling=z["S"]), z["X_rec_s"], rtol=1e-10)
    np.testing.assert_allclose(spr.unscale_data(z["x0"], sampling=z["S"]), z["x_uns"], rtol=1e-13)


@pytest.mark.parametrize("n_sensors,d_min", [(10, 0.0), (16, 0.06)])
def test_gem_row_sharded_matches_single_rank(torch_cuda, n_sensors, d_min):
    """GEM over a row-sharded basis (thread-emulated ranks): every rank offers its local winner, the largest value
    (lowest global index on ties) wins -- the selection must be the single-rank one, on every rank."""
    import threading
    torch = torch_cuda
    from oracle import synth as osynth
    from openmeasure_b200 import comm as C, engine as E
    F, m, r = 3, 40, 20
    cells = [900, 700, 1100]
    n_c = sum(cells)
    X = osynth.snapshots(F, n_c, m, r)
    rng = np.random.default_rng(11)
    xyz = rng.random((n_c, 3))
    mask = rng.random(F * n_c) > 0.15
    one = _sps().SPR(X, F, xyz)
    one.fit(select_modes="number", n_modes=r)
    Ur = one.Ur
    draws = [rng.standard_normal(k) for k in range(n_sensors + 1)]
    take = lambda seq: (lambda size: seq.pop(0)[:size])
    ref = one._eng.gem(n_sensors, torch.from_numpy(mask).cuda(), torch.from_numpy(xyz).cuda(), d_min,
                       normal=take([d.copy() for d in draws[2:]]))
    comms = C.ThreadComm.make(len(cells))
    out, err = [None] * len(cells), []

    def run(rk):
        try:
            lay = C.ShardLayout(F, cells, rk)
            rows = lay.to_global(torch.arange(F * cells[rk])).numpy()
            eng = E.Engine(torch.from_numpy(np.ascontiguousarray(X[rows])).cuda(), F, comm=comms[rk])
            eng.set_basis_rows(torch.from_numpy(np.ascontiguousarray(Ur[rows])).cuda())
            xl = xyz[lay.cell0:lay.cell0 + cells[rk]]
            out[rk] = eng.gem(n_sensors, torch.from_numpy(mask[rows]).cuda(), torch.from_numpy(np.ascontiguousarray(xl)).cuda(),
                              d_min, normal=take([d.copy() for d in draws[2:]]))
        except Exception as e:          # pragma: no cover
            err.append(e)
            comms[rk].shared.barrier.abort()

    th = [threading.Thread(target=run, args=(k,)) for k in range(len(cells))]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    for o in out:
        np.testing.assert_array_equal(o, ref)


# ---------------------------------------------------------------------------------------------
# SURVEY 8f rows 2 and 3 on repo kernels: weighted OLS (batched Householder QR) and general CSR measurement matrices
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("s,r,N", [(12, 12, 5), (40, 24, 33), (100, 100, 64), (150, 90, 10)])
def test_weighted_ols_kernel_matches_pinv_loop(torch_cuda, s, r, N):
    """SPR.predict's weighted branch (sparse_sensing.py:871-878) per vector with np.linalg.pinv, against the batched
    QR kernel; one vector gets a rank-deficient weighted Theta and must come back through the pseudo-inverse route."""
    torch = torch_cuda
    from openmeasure_b200 import engine as E
    rng = np.random.default_rng(s * r + N)
    Theta = rng.standard_normal((s, r))
    y0v = rng.standard_normal((N, s))
    y0s = 0.05 + rng.random((N, s))
    eng = E.Engine(torch.zeros(8, 4, dtype=torch.float64, device="cuda"), 1, group=False)
    a, sg = eng.wols_predict(torch.from_numpy(Theta).cuda(), torch.from_numpy(y0v).cuda(), torch.from_numpy(y0s).cuda())
    a, sg = a.cpu().numpy(), sg.cpu().numpy()
    for b in range(N):
        W = np.diag(1 / y0s[b])
        P = np.linalg.pinv(W @ Theta)
        np.testing.assert_allclose(a[b], P @ (W @ y0v[b]), rtol=1e-9, atol=1e-9 * np.abs(a[b]).max())
        np.testing.assert_allclose(sg[b], np.abs(P @ y0s[b]), rtol=1e-9, atol=1e-9 * np.abs(sg[b]).max())
    Th2 = Theta.copy()
    Th2[:, -1] = Th2[:, 0]                                             # rank deficient: the kernel must flag it
    a2, _ = eng.wols_predict(torch.from_numpy(Th2).cuda(), torch.from_numpy(y0v[:2]).cuda(), torch.from_numpy(y0s[:2]).cuda())
    for b in range(2):
        W = np.diag(1 / y0s[b])
        np.testing.assert_allclose(a2[b].cpu().numpy(), np.linalg.pinv(W @ Th2, rcond=1e-15) @ (W @ y0v[b]), rtol=1e-6, atol=1e-8)


def test_general_csr_matrix_never_densified_and_row_sharded(torch_cuda, monkeypatch):
    """train(C) / reconstruct(sampling=) with a scipy CSR line-of-sight style matrix: equals the dense products of the
    reference (sparse_sensing.py:797, :573, :365-368), never calls toarray(), and gives the same Theta when the rows
    are sharded over (thread-emulated) ranks."""
    import threading
    import scipy.sparse as sp
    torch = torch_cuda
    from oracle import pod_oracle as po, synth as osynth
    from openmeasure_b200 import comm as Cm
    F, m, r, s = 3, 32, 10, 17
    cells = [500, 420, 380]
    n_c = sum(cells)
    n = F * n_c
    X = osynth.snapshots(F, n_c, m, r)
    rng = np.random.default_rng(5)
    Cs = sp.random(s, n, density=0.01, format="csr", random_state=3, data_rvs=lambda k: rng.random(k) + 0.1)
    Cs[0, :] = 0                                                       # an empty row
    Cs.eliminate_zeros()
    dense = Cs.toarray()
    monkeypatch.setattr(sp.csr_matrix, "toarray", lambda *a, **k: (_ for _ in ()).throw(AssertionError("densified")))
    monkeypatch.setattr(sp.csr_matrix, "todense", lambda *a, **k: (_ for _ in ()).throw(AssertionError("densified")))
    one = _sps().SPR(X, F, np.zeros((n_c, 3)))
    one.fit(select_modes='number', n_modes=r)
    one.train(Cs, cond=True)
    Ur = one.Ur
    np.testing.assert_allclose(one.Theta, dense @ Ur, rtol=0, atol=1e-12 * np.abs(dense @ Ur).max())
    np.testing.assert_allclose(one._cnt_s.cpu().numpy(), dense @ one.X_cnt[:, 0], rtol=1e-12)
    A = rng.standard_normal((4, r))
    np.testing.assert_allclose(one.reconstruct(A, sampling=Cs),
                               (dense @ one.X_scl[:, 0])[:, None] * ((dense @ Ur) @ A.T) + (dense @ one.X_cnt[:, 0])[:, None],
                               rtol=1e-10, atol=1e-10)
    comms = Cm.ThreadComm.make(len(cells))
    out, err = [None] * len(cells), []

    def run(rk):
        try:
            lay = Cm.ShardLayout(F, cells, rk)
            rows = lay.to_global(torch.arange(F * cells[rk])).numpy()
            spr = _sps().SPR.from_device(torch.from_numpy(np.ascontiguousarray(X[rows])).cuda(), F, comm=comms[rk])
            spr.fit(select_modes='number', n_modes=r)
            spr.train(Cs)
            out[rk] = (spr.Theta.copy(), spr._cnt_s.cpu().numpy().copy())
        except Exception as e:          # pragma: no cover
            err.append(e)
            comms[rk].shared.barrier.abort()

    th = [threading.Thread(target=run, args=(k,)) for k in range(len(cells))]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0]
    for Th, cs in out:
        np.testing.assert_allclose(np.abs(Th), np.abs(one.Theta), rtol=0, atol=1e-10 * np.abs(one.Theta).max())
        np.testing.assert_allclose(cs, one._cnt_s.cpu().numpy(), rtol=1e-12)
    np.testing.assert_array_equal(out[0][0], out[1][0])                # identical bits on every rank


def test_bad_sensor_and_feature_indices_raise_index_error(torch_cuda):
    from oracle import synth as osynth
    F, n_c, m, r = 2, 300, 16, 6
    X = osynth.snapshots(F, n_c, m, r)
    spr = _sps().SPR(X, F, np.zeros((n_c, 3)))
    spr.fit(select_modes='number', n_modes=r)
    C = spr.optimal_placement()
    bad = _sps().SensorMatrix(np.array([0, 5, F * n_c + 3, 7, 8, 9]), F * n_c)
    with pytest.raises(IndexError):
        spr.train(bad)
    spr.train(C)
    y = np.zeros((r, 3))
    y[:, 2] = F                                                        # feature id out of range
    with pytest.raises(IndexError):
        spr.predict(y)


# ---------------------------------------------------------------------------------------------
# streamed reconstruct (configs[3]: the n x N result never exists on the device): row chunks through the
# two-buffer ring must equal the one-shot result bit for bit, for every `out` form
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_c,r,N,chunk", [(700, 10, 7, 128), (1100, 16, 16, 256), (3001, 40, 130, 1024), (900, 14, 258, 384)])
def test_reconstruct_streams_in_row_chunks(torch_cuda, n_c, r, N, chunk):
    from oracle import pod_oracle as po, synth as osynth
    F, m = 3, max(24, r + 8)
    X = osynth.snapshots(F, n_c, m, r)
    ref = po.fit(X, F, "std", 1, "number", r)
    spr = _sps().SPR(X, F, np.zeros((n_c, 3)))
    spr.fit(select_modes='number', n_modes=r)
    rng = np.random.default_rng(N)
    A = rng.standard_normal((N, r))
    whole = spr.reconstruct(A, chunk_rows=1 << 30)                      # one chunk
    Ur, sg = _sign_align(spr.Ur, ref["Ur"])
    np.testing.assert_allclose(Ur, ref["Ur"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(whole, po.reconstruct(ref["Ur"], A * sg, ref["X_cnt"], ref["X_scl"]), rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(spr.reconstruct(A, chunk_rows=chunk), whole)
    buf = np.full((F * n_c, N), np.nan)
    assert spr.reconstruct(A, out=buf, chunk_rows=chunk) is buf
    np.testing.assert_array_equal(buf, whole)
    seen = []
    got = np.empty_like(whole)

    def sink(row0, block):
        seen.append((row0, block.shape[0]))
        got[row0:row0 + block.shape[0]] = block

    assert spr.reconstruct(A, out=sink, chunk_rows=chunk) is None
    np.testing.assert_array_equal(got, whole)
    assert [s0 for s0, _ in seen] == list(range(0, F * n_c, chunk)) and sum(k for _, k in seen) == F * n_c
    blocks = [(r0, b.copy()) for r0, b in spr.reconstruct_chunks(A[:1], chunk_rows=chunk)]
    np.testing.assert_array_equal(np.vstack([b for _, b in blocks]), whole[:, :1])
    with pytest.raises(ValueError):
        spr.reconstruct(A, out=np.empty((3, 3)))


def test_from_host_shard_matches_constructor(torch_cuda):
    """SPR.from_host (one rank's pinned host shard, block-wise upload overlapped with the first passes) against
    the plain constructor on the same matrix."""
    torch = torch_cuda
    from oracle import synth as osynth
    F, n_c, m, r = 4, 2500, 41, 12
    X = osynth.snapshots(F, n_c, m, r)
    Xp = torch.from_numpy(X).pin_memory().numpy()
    a = _sps().SPR(X, F, np.zeros((n_c, 3)))
    b = _sps().SPR.from_host(Xp, F, np.zeros((n_c, 3)), group=False)
    for s in (a, b):
        s.fit(select_modes='number', n_modes=r)
        s.optimal_placement()
    np.testing.assert_array_equal(a.X_cnt, b.X_cnt)
    np.testing.assert_array_equal(a.X_scl, b.X_scl)
    np.testing.assert_allclose(a.Sigma_r, b.Sigma_r, rtol=1e-13)
    np.testing.assert_array_equal(a.qr_pivots, b.qr_pivots)
    with pytest.raises(TypeError):
        _sps().SPR.from_host(torch.from_numpy(X), F)


# ---------------------------------------------------------------------------------------------
# BASELINE shapes: configs[0] (README 2D ROM: 9 x 18 362 x 41, r = 14) end to end against the oracle, and
# configs[1] pivots against LAPACK run on the GPU's own basis
# ---------------------------------------------------------------------------------------------
def test_config1_shape_full_pipeline_matches_oracle(torch_cuda):
    from oracle import pod_oracle as po, synth as osynth
    F, n_c, m, r = 9, 18362, 41, 14
    X = osynth.snapshots(F, n_c, m, r)
    ref = po.placement_pipeline(X, F, r)
    spr = _sps().SPR(X, F, np.zeros((n_c, 3)))
    spr.fit(select_modes='number', n_modes=r)
    C = spr.optimal_placement()
    np.testing.assert_array_equal(spr.X_cnt, ref["X_cnt"])
    np.testing.assert_array_equal(spr.X_scl, ref["X_scl"])
    np.testing.assert_allclose(spr.Sigma_r, ref["Sigma_r"], rtol=RTOL)
    Ur, sg = _sign_align(spr.Ur, ref["Ur"])
    np.testing.assert_allclose(Ur, ref["Ur"], rtol=0, atol=1e-10)
    np.testing.assert_array_equal(spr.qr_pivots, ref["piv"])
    spr.train(C)
    Co = po.one_hot(ref["piv"], X.shape[0])
    Th = po.theta(Co, ref["Ur"])
    np.testing.assert_allclose(spr.Theta * sg, Th, rtol=0, atol=1e-10)
    ys = []
    for j in (0, 7, 40):
        y = np.zeros((r, 3))
        y[:, 0] = X[ref["piv"], j]
        y[:, 2] = ref["piv"] // n_c
        ys.append(y)
    a, _ = spr.predict(ys)
    ao, _ = po.predict_ols(Th, ys, Co, ref["X_cnt"], ref["X_scl"], n_c)
    np.testing.assert_allclose(a * sg, ao, rtol=1e-9, atol=1e-9 * np.abs(ao).max())
    np.testing.assert_allclose(spr.reconstruct(a), po.reconstruct(ref["Ur"], ao, ref["X_cnt"], ref["X_scl"]), rtol=RTOL)


def test_config2_shape_pivots_match_lapack_on_device_basis(torch_cuda):
    """configs[1] (1 652 580 x 41, r = 40): the blocked GPU placement against scipy.linalg.qr(pivoting=True) run on
    the basis the GPU produced (the reference's own call, sparse_sensing.py:739; ~10 s of LAPACK)."""
    torch = torch_cuda
    from openmeasure_b200 import synth as gsynth
    F, n_c, m, r = 9, 183620, 41, 40
    spr = _sps().SPR.from_device(gsynth.snapshots(F, n_c, m, r), F, group=False)
    spr.fit(select_modes='number', n_modes=r)
    spr.optimal_placement()
    Ur = spr.Ur
    _, _, P = sla.qr(Ur.T, pivoting=True, mode='economic')
    np.testing.assert_array_equal(spr.qr_pivots, P[:r])


# ---------------------------------------------------------------------------------------------
# one process, several devices, the UNCHANGED constructor (ROM.devices / OMB_DEVICES): here three "devices" that are
# all cuda:0 (the driver's box has one GPU); tools/multi_device_check.py runs the same on real GPUs
# ---------------------------------------------------------------------------------------------
def test_unchanged_constructor_drives_several_devices(torch_cuda, monkeypatch):
    from oracle import pod_oracle as po, synth as osynth
    sps = _sps()
    F, n_c, m, r = 3, 2300, 32, 12
    X = osynth.snapshots(F, n_c, m, r)
    xyz = np.random.default_rng(0).random((n_c, 3))
    ref = po.placement_pipeline(X, F, r)
    one = sps.SPR(X, F, xyz)
    one.fit(select_modes='number', n_modes=r)
    C1 = one.optimal_placement()
    one.train(C1)
    monkeypatch.setattr(sps.ROM, "devices", [0, 0, 0])
    spr = sps.SPR(X, F, xyz)                                   # the reference's constructor, nothing else
    assert type(spr).__name__ == "_MultiSPR" and isinstance(spr, sps.SPR) and spr._md.G == 3
    spr.fit(select_modes='number', n_modes=r)
    C = spr.optimal_placement()
    np.testing.assert_array_equal(spr.X_cnt, ref["X_cnt"])
    np.testing.assert_allclose(spr.X_scl, ref["X_scl"], rtol=1e-15)
    np.testing.assert_allclose(spr.Sigma_r, ref["Sigma_r"], rtol=RTOL)
    np.testing.assert_array_equal(C.pivots, ref["piv"])
    Ur, _ = _sign_align(spr.Ur, ref["Ur"])
    np.testing.assert_allclose(Ur, ref["Ur"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(spr.X0, ref["X0"] if "X0" in ref else one.X0, rtol=1e-13, atol=1e-13)
    spr.train(C)
    np.testing.assert_allclose(np.abs(spr.Theta), np.abs(one.Theta), rtol=0, atol=1e-10)
    ys = []
    for j in (1, 9):
        y = np.zeros((r, 3))
        y[:, 0] = X[ref["piv"], j]
        y[:, 2] = ref["piv"] // n_c
        ys.append(y)
    a, _ = spr.predict(ys)
    a1, _ = one.predict(ys)
    np.testing.assert_allclose(np.abs(a), np.abs(a1), rtol=1e-8, atol=1e-9 * np.abs(a1).max())
    rec, rec1 = spr.reconstruct(a, chunk_rows=256), one.reconstruct(a1)
    np.testing.assert_allclose(rec, rec1, rtol=1e-9)
    blocks = sorted(((g0, b.copy()) for g0, b in spr.reconstruct_chunks(a, chunk_rows=512)), key=lambda t: t[0])
    np.testing.assert_allclose(np.vstack([b for _, b in blocks]), rec1, rtol=1e-9)
    x0 = np.random.default_rng(1).standard_normal(F * n_c)
    np.testing.assert_allclose(spr.unscale_data(x0), one.unscale_data(x0), rtol=1e-14)
    mask = np.random.default_rng(2).random(F * n_c) > 0.3
    Cg = spr.optimal_placement(calc_type='gem', n_sensors=6, mask=mask)
    np.random.seed(0)
    assert Cg.shape == (6, F * n_c) and mask[Cg.pivots].all()
    with pytest.raises(ValueError):
        spr.fit(select_modes='bogus')


def test_config3_snapshot_and_sensor_counts_match_oracle(torch_cuda):
    """configs[2]'s m = 256 snapshots and r = s = 100 sensors on 1/90 of its rows (9 x 20 000 cells): every stage of the
    many-snapshot route against the oracle (= the reference's np.linalg.svd and scipy.linalg.qr calls, ~10 s of CPU)."""
    from oracle import pod_oracle as po, synth as osynth
    F, n_c, m, r = 9, 20000, 256, 100
    X = osynth.snapshots(F, n_c, m, r)
    ref = po.placement_pipeline(X, F, r)
    spr = _sps().SPR(X, F, np.zeros((n_c, 3)))
    spr.fit(select_modes='number', n_modes=r)
    C = spr.optimal_placement()
    np.testing.assert_array_equal(spr.X_cnt, ref["X_cnt"])
    np.testing.assert_array_equal(spr.X_scl, ref["X_scl"])
    np.testing.assert_allclose(spr.Sigma_r, ref["Sigma_r"], rtol=RTOL)
    np.testing.assert_array_equal(spr.qr_pivots, ref["piv"])
    Ur, sg = _sign_align(spr.Ur, ref["Ur"])
    np.testing.assert_allclose(Ur[:, :r // 2], ref["Ur"][:, :r // 2], rtol=0, atol=1e-9)
    spr.train(C)
    ys = []
    for j in (0, 100, 255):
        y = np.zeros((r, 3))
        y[:, 0] = X[ref["piv"], j]
        y[:, 2] = ref["piv"] // n_c
        ys.append(y)
    a, _ = spr.predict(ys)
    Co = po.one_hot(ref["piv"], X.shape[0])
    ao, _ = po.predict_ols(po.theta(Co, ref["Ur"]), ys, Co, ref["X_cnt"], ref["X_scl"], n_c)
    np.testing.assert_allclose(spr.reconstruct(a), po.reconstruct(ref["Ur"], ao, ref["X_cnt"], ref["X_scl"]), rtol=1e-9)
