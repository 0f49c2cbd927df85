"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

/root/reference does not exist on the GPU box, so the outputs are committed as small fixtures.
The reference module imports cvxpy at module scope (sparse_sensing.py:15) but the hot path only
touches cp.multiply(a, b) + c -> .value (sparse_sensing.py:233-238); cvxpy is not installed here,
so a minimal stand-in exposing just that is put on sys.path (it changes no reference code).
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def _import_reference():
    try:
        import cvxpy  # noqa: F401
    except ImportError:
        shim = types.ModuleType("cvxpy")

        class _Expr:
            def __init__(self, v):
                self.value = np.asarray(v)

            def __add__(self, o):
                return _Expr(self.value + (o.value if isinstance(o, _Expr) else o))

            __radd__ = __add__

        shim.multiply = lambda a, b: _Expr(np.asarray(a) * np.asarray(b))
        sys.modules["cvxpy"] = shim
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.join(REF, "src"))
    import openmeasure.sparse_sensing as sps
    return sps


def _case(sps, name, X, F, n_modes, select_modes="number", scale_type="std", axis_cnt=1,
          n_meas=3, sigma=0.0, rng=None):
    n, m = X.shape
    n_c = n // F
    xyz = np.zeros((n_c, 3))
    spr = sps.SPR(X.copy(), F, xyz)
    spr.fit(scale_type=scale_type, axis_cnt=axis_cnt, select_modes=select_modes, n_modes=n_modes)
    Ur0 = spr.Ur.copy()
    C = spr.optimal_placement()
    piv = np.argmax(C, axis=1).astype(np.int64)
    spr.train(C)
    ys = []
    for t in range(n_meas):
        x_new = X[:, t % m] * (1.0 + 0.01 * rng.standard_normal(n))
        y = np.zeros((len(piv), 3))
        y[:, 0] = x_new[piv]
        y[:, 1] = sigma * np.abs(y[:, 0]) if sigma else 0.0
        y[:, 2] = piv // n_c
        ys.append(y)
    Ar_p, Ar_sig = spr.predict(ys)
    Xrec = spr.reconstruct(Ar_p)
    out = dict(
        X=X, F=np.int64(F), n_modes=np.float64(n_modes), select_modes=select_modes,
        scale_type=scale_type, axis_cnt=np.int64(-1 if axis_cnt is None else axis_cnt),
        X_cnt=spr.X_cnt, X_scl=spr.X_scl, r=np.int64(spr.r), Sigma_r=spr.Sigma_r,
        Ur=Ur0, Ar=spr.Ar, Vr=spr.Vr, piv=piv, Theta=spr.Theta,
        Y=np.stack(ys), Ar_pred=Ar_p, Ar_sigma=Ar_sig, X_rec=Xrec,
        X0_checksum=np.array([spr.X0.sum(), np.abs(spr.X0).sum()]),
        X0_head=spr.X0[: min(n, 64)].copy(),
    )
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} m={m} r={spr.r} piv[:6]={piv[:6]} -> {os.path.getsize(path)/1e3:.0f} kB")


def _gem_case(sps, name, X, F, n_modes, n_sensors, d_min, seed, rng):
    """optimal_placement(calc_type='gem') of the unmodified reference; its random jitter
    (sparse_sensing.py:667) is made reproducible by seeding numpy's global generator."""
    n, m = X.shape
    n_c = n // F
    xyz = rng.random((n_c, 3))
    mask = rng.random(n) > 0.1
    spr = sps.SPR(X.copy(), F, xyz)
    spr.fit(select_modes="number", n_modes=n_modes)
    np.random.seed(seed)
    C = spr.optimal_placement(calc_type="gem", n_sensors=n_sensors, mask=mask, d_min=d_min)
    piv = np.argmax(C, axis=1).astype(np.int64)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, X=X, F=np.int64(F), n_modes=np.int64(n_modes), xyz=xyz, mask=mask, Ur=spr.Ur,
                        n_sensors=np.int64(n_sensors), d_min=np.float64(d_min), seed=np.int64(seed), gem=piv)
    print(f"{name}: n={n} r={spr.r} gem sensors={piv} -> {os.path.getsize(path)/1e3:.0f} kB")


def _scaling_case(sps, name, rng):
    """ROM.scale_data of the unmodified reference for EVERY scaling type it implements
    (sparse_sensing.py:114-161), both centring modes: X_cnt, X_scl and X0 on one small matrix."""
    F, n_c, m = 3, 150, 11
    X = rng.random((F * n_c, m)) * 10.0 ** rng.integers(-1, 3, (F * n_c, 1)) + 0.5
    out = dict(X=X, F=np.int64(F))
    for st in ("std", "none", "pareto", "vast", "range", "level", "max", "variance", "median", "poisson", "l2-norm"):
        for ax, tag in ((1, "row"), (None, "blk")):
            rom = sps.ROM(X.copy(), F, None)
            X0 = rom.scale_data(st, ax)
            key = st.replace("-", "_") + "_" + tag
            out[key + "_cnt"], out[key + "_scl"], out[key + "_X0"] = rom.X_cnt, rom.X_scl, X0
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {len(out) - 2} arrays -> {os.path.getsize(path)/1e3:.0f} kB")


def _extras_case(sps, name, rng):
    """The API corners around the path, from the unmodified reference: fit(basis=...), train with a general
    (non one-hot) dense C and cond=True (:797, :813-820), weighted predict (:871-878), reconstruct and
    unscale_data with a sampling matrix (:365-368, :233)."""
    from oracle import synth
    F, n_c, m, r = 3, 300, 20, 8
    X = synth.snapshots(F, n_c, m, r)
    n = F * n_c
    spr = sps.SPR(X.copy(), F, np.zeros((n_c, 3)))
    spr.fit(select_modes="number", n_modes=r)
    Ur, Ar_fit = spr.Ur.copy(), spr.Ar.copy()
    s = 12
    C = np.zeros((s, n))
    for i in range(s):                                   # line-of-sight-like rows: a few cells of one feature
        f = i % F
        cols = f * n_c + rng.choice(n_c, 5, replace=False)
        C[i, cols] = rng.random(5)
    spr.train(C, cond=True)
    ys = []
    for t in range(3):
        x = X[:, t] * (1.0 + 0.01 * rng.standard_normal(n))
        y = np.zeros((s, 3))
        y[:, 0] = C @ x
        y[:, 1] = 0.02 * np.abs(y[:, 0]) + 1e-3
        y[:, 2] = np.arange(s) % F
        ys.append(y)
    Ar_p, Ar_sig = spr.predict(ys)
    S = rng.random((9, n)) * (rng.random((9, n)) < 0.02)
    Xs = spr.reconstruct(Ar_p, sampling=S)
    x0 = rng.standard_normal(9)
    xs = spr.unscale_data(x0, sampling=S)
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, X=X, F=np.int64(F), r=np.int64(r), Ur=Ur, Ar=Ar_fit, C=C, Theta=spr.Theta, k=np.float64(spr.k),
                        Y=np.stack(ys), Ar_pred=Ar_p, Ar_sigma=Ar_sig, S=S, X_rec_s=Xs, x0=x0, x_uns=xs,
                        cnt_vector=spr.cnt_vector, scl_vector=spr.scl_vector)
    print(f"{name}: k={spr.k:.3f} -> {os.path.getsize(path)/1e3:.0f} kB")


def main():
    sps = _import_reference()
    sys.path.insert(0, ROOT)
    from oracle import synth
    os.makedirs(GOLD, exist_ok=True)
    rng = np.random.default_rng(20261018)
    partial = bool({"--gem-only", "--scaling", "--extras"} & set(sys.argv))   # leave the g1..g5 fixtures untouched
    if not partial:

        # g1: the reference unit tests' own shape (tests/test_rom.py:8-13: 2 features x 10 points x 5)
        _case(sps, "g1_unit_20x5", rng.random((20, 5)), 2, 4, rng=rng)
        # g2: random 3 features x 400 cells x 12 snapshots, variance-based rank, pareto scaling
        _case(sps, "g2_rand_1200x12_pareto", rng.random((1200, 12)) + 0.5, 3, 95.0,
              select_modes="variance", scale_type="pareto", rng=rng)
        # g3: synthetic generator, README-like layout (9 features), scaled down: 9 x 400 x 41, r = 14
        _case(sps, "g3_synth_3600x41_r14", synth.snapshots(9, 400, 41, 14), 9, 14, rng=rng)
        # g4: range scaling, scalar centring (axis_cnt=None), weighted measurements (sigma != 0)
        _case(sps, "g4_synth_2400x24_range", synth.snapshots(4, 600, 24, 10), 4, 10,
              scale_type="range", axis_cnt=None, sigma=0.02, rng=rng)
        # g5: wider snapshot set, m > 128 exercises the multi-leaf row-mean tree: 2 x 300 x 160, r = 20
        _case(sps, "g5_synth_600x160_r20", synth.snapshots(2, 300, 160, 20), 2, 20, rng=rng)
    # s1: every scaling type x both centring modes (never overwritten unless --scaling is given)
    if "--scaling" in sys.argv or not os.path.exists(os.path.join(GOLD, "s1_scalings_450x11.npz")):
        _scaling_case(sps, "s1_scalings_450x11", np.random.default_rng(99))
    if "--scaling" in sys.argv:
        return
    # x1: general C + cond, weighted predict, sampling matrices
    if "--extras" in sys.argv or not os.path.exists(os.path.join(GOLD, "x1_extras_900x20_r8.npz")):
        _extras_case(sps, "x1_extras_900x20_r8", np.random.default_rng(5))
    if "--extras" in sys.argv:
        return
    # g6/g7: GEM placement (sparse_sensing.py:586-698) with a user mask, without / with a d_min radius
    rng2 = np.random.default_rng(7)
    _gem_case(sps, "g6_gem_900x20_r8", synth.snapshots(3, 300, 20, 8), 3, 8, 6, 0.0, 11, rng2)
    _gem_case(sps, "g7_gem_1200x24_r10_dmin", synth.snapshots(2, 600, 24, 10), 2, 10, 9, 0.08, 12, rng2)


if __name__ == "__main__":
    main()
