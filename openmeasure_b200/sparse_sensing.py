"""B200-native mirror of OpenMEASURE's snapshot-POD sparse-sensing classes.

Drop-in for the hot path of /root/reference/src/openmeasure/sparse_sensing.py:

    ROM/SPR(X, n_features, xyz) -> fit(...) -> optimal_placement() -> train(C) -> predict(y)
    -> reconstruct(ap)

Same class and method names, argument meaning, attributes and exception types as the reference
(ROM: sparse_sensing.py:18-511, SPR: :513-901); the numerics run on the GPU through the C ABI of
libomb200.so (see engine.py).  There is no CPU fallback: without the CUDA library and a CUDA
device every compute call raises.

Deliberate differences (DESIGN.md "Reference quirks"):
  * optimal_placement() returns a lazy one-hot `SensorMatrix` (shape, C[i, :], C @ x, C.dot,
    np.asarray(C) all work) instead of a dense s x n float array (13 GB at config 3);
  * X0 and Ur are materialised on the host only when read;
  * POD modes are defined up to sign (as in any SVD); singular values/reconstructions agree with
    the reference to 1e-10, pivots are identical on non-degenerate inputs.
Out of scope (raise NotImplementedError): constrained OLS ('COLS'), CPOD, adaptive_sampling,
scale_limits -- see SURVEY.md section 8.
"""
import numpy as np
import torch

from . import engine as _eng

_BROKEN_SCALES = ("vast_2", "vast_3", "vast_4")   # broadcast-fail in the reference (:147-157)


class SensorMatrix:
    """One-hot measurement matrix C (s x n): row j has a single 1 at column pivots[j].

    Stands in for the dense array built at sparse_sensing.py:740-743.  Supports the idioms the
    reference's README and tests use: C.shape, C[i, :], np.argmax(C[i, :]), C @ x, C.dot(x),
    np.asarray(C)."""

    def __init__(self, pivots, n):
        self.pivots = np.asarray(pivots, dtype=np.int64)
        self.shape = (int(self.pivots.size), int(n))
        self.ndim = 2
        self.dtype = np.dtype(np.float64)

    def toarray(self):
        C = np.zeros(self.shape)
        C[np.arange(self.shape[0]), self.pivots] = 1
        return C

    def __array__(self, dtype=None, copy=None):
        a = self.toarray()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2:
            rows, cols = key
        else:
            rows, cols = key, slice(None)
        ridx = np.arange(self.shape[0])[rows]
        if np.ndim(ridx) == 0:
            row = np.zeros(self.shape[1])
            row[self.pivots[ridx]] = 1
            return row[cols]
        sub = np.zeros((len(ridx), self.shape[1]))
        sub[np.arange(len(ridx)), self.pivots[ridx]] = 1
        return sub[:, cols]

    def dot(self, x):
        x = np.asarray(x)
        if x.shape[0] != self.shape[1]:
            raise ValueError("shapes not aligned")
        return x[self.pivots]

    __matmul__ = dot


def _check_rows(piv, n):
    """Row indices of a one-hot matrix must address rows of X (the reference's dense C cannot hold anything else)."""
    piv = np.asarray(piv)
    if piv.size and (piv.min() < 0 or piv.max() >= n):
        raise IndexError('sensor index out of range: C has a 1 outside the %d rows of X' % n)


def _one_hot_csr(piv, n):
    import scipy.sparse as sp
    piv = np.asarray(piv, dtype=np.int64)
    return sp.csr_matrix((np.ones(piv.size), (np.arange(piv.size), piv)), shape=(piv.size, n))


def _as_pivots(C):
    """pivots of a one-hot C (SensorMatrix or dense), else None."""
    if isinstance(C, SensorMatrix):
        return C.pivots
    if isinstance(C, np.ndarray) and C.ndim == 2:
        piv = np.argmax(C, axis=1)
        ok = (C[np.arange(C.shape[0]), piv] == 1).all() and (np.count_nonzero(C, axis=1) == 1).all()
        return piv.astype(np.int64) if ok else None
    return None


def _upload_blocks(X, n_features):
    """Host snapshot matrix -> HBM.  A pinned matrix is uploaded one feature block at a time on a copy
    stream and the per-block arrival events are returned: the engine's first stage consumes the blocks
    as they land (Engine.stats), so the statistics / Gram passes overlap the PCIe transfer instead of
    following it.  Pageable memory: one synchronous copy, arrival = None."""
    if X.dtype != np.float64 or not X.flags.c_contiguous:
        X = np.ascontiguousarray(X, dtype=np.float64)
    dev = torch.device("cuda", torch.cuda.current_device())
    Xd = torch.empty(X.shape, dtype=torch.float64, device=dev)
    Xh = torch.from_numpy(X)
    F, n_c = n_features, X.shape[0] // n_features
    if not (Xh.is_pinned() and F > 1):
        Xd.copy_(Xh, non_blocking=True)
        return Xd, None
    copy_stream = torch.cuda.Stream(device=dev)
    copy_stream.wait_stream(torch.cuda.current_stream(dev))
    arrival = []
    with torch.cuda.stream(copy_stream):
        for f in range(F):
            Xd[f * n_c:(f + 1) * n_c].copy_(Xh[f * n_c:(f + 1) * n_c], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            arrival.append(ev)
    Xd.record_stream(copy_stream)
    return Xd, arrival


class ROM:
    """Reduced-order-model utilities: centring/scaling, POD, truncation, reconstruction
    (reference ROM, sparse_sensing.py:18-511)."""

    # POD accuracy control (extension): the Gram route resolves sigma_r to eps (sigma_1/sigma_r)^2;
    # when that estimate exceeds pod_refine_tol a CholeskyQR2-style correction of the basis (two
    # more passes over the n x r modes) brings it to eps sigma_1/sigma_r; when even that is not
    # enough (sigma_r/sigma_1 < ~1e-5, or a retained mode below sqrt(m eps) sigma_1 that the Gram
    # matrix cannot hold at all) the full-width route (_pod_full_width: the accuracy class of a TSQR)
    # runs.  'auto' | True (always correct) | 'full' (always full width) | False.
    pod_refine = 'auto'
    pod_refine_tol = 1e-11
    # Extension: devices one process drives behind this very constructor.  None (or one id): the current CUDA
    # device.  'all' or a list of ids (or OMB_DEVICES=all | 0,1,.. in the environment): the rows are sharded over
    # those GPUs, one host thread each (openmeasure_b200/multi.py); every method keeps its reference signature.
    devices = None

    def __new__(cls, *args, **kwargs):
        if cls is ROM or cls is SPR:
            from . import multi as _multi
            if len(_multi.resolve_devices(cls.devices)) > 1:
                return object.__new__(_MultiSPR if cls is SPR else _MultiROM)
        return object.__new__(cls)

    def __init__(self, X, n_features, xyz):
        if type(X) is not np.ndarray:                     # :69-70
            raise TypeError('The matrix X is not a numpy array.')
        if type(n_features) is not int:                   # :71-72
            raise TypeError('The parameter n_features is not an integer.')
        self.X = X
        self.n_features = n_features
        self.xyz = xyz
        n = X.shape[0]
        self.n_points = n // n_features
        if n % n_features != 0:                           # :80-81
            raise Exception('The number of rows of X is not a multiple of n_features')
        self._eng = None
        self._host = {}

    # ------------------------------------------------------------------ device plumbing
    @classmethod
    def from_device(cls, X_dev, n_features, xyz=None, group=None, comm=None):
        """Extension: build on an HBM-resident shard (torch CUDA float64 (F*n_c_loc, m)).  With
        torch.distributed initialised (one process per GPU) each rank passes its own cells
        [c0, c0 + n_c_loc) of every feature; pivots / C then refer to GLOBAL row indices
        f * n_c + c, Theta / predict are replicated, reconstruct returns the local rows."""
        self = object.__new__(cls)
        if type(n_features) is not int:
            raise TypeError('The parameter n_features is not an integer.')
        if X_dev.shape[0] % n_features != 0:
            raise Exception('The number of rows of X is not a multiple of n_features')
        self.X = None
        self.n_features = n_features
        self.xyz = xyz
        self._eng = _eng.Engine(X_dev, n_features, group=group, comm=comm)
        self.n_points = self._eng.n_c_loc
        self._host = {}
        return self

    @classmethod
    def from_npy(cls, path, n_features, xyz=None, group=None):
        """Extension: build from an on-disk .npy snapshot matrix (README.md:50-53), streamed straight
        into HBM (openmeasure_b200/ingest.py).  One process per GPU: every rank reads only its own
        cells of every feature; `xyz` are then the coordinates of those cells."""
        from . import ingest
        rank, world = 0, 1
        if group is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
            rank, world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
        if type(n_features) is not int:
            raise TypeError('The parameter n_features is not an integer.')
        Xd = ingest.load_npy_shard(path, n_features, rank, world)
        return cls.from_device(Xd, n_features, xyz, group=group)

    @classmethod
    def from_host(cls, X_host, n_features, xyz=None, group=None):
        """Extension: one rank's HOST shard of a row-sharded problem (numpy float64 (F*n_c_loc, m): this
        rank's cells [c0, c0 + n_c_loc) of every feature; pinned memory makes the upload asynchronous).
        The same upload path as the plain constructor -- feature blocks cross PCIe on a copy stream while
        the statistics / Gram passes consume the blocks that have landed -- followed by from_device()."""
        if type(X_host) is not np.ndarray:
            raise TypeError('The matrix X is not a numpy array.')
        if type(n_features) is not int:
            raise TypeError('The parameter n_features is not an integer.')
        if X_host.shape[0] % n_features != 0:
            raise Exception('The number of rows of X is not a multiple of n_features')
        _eng.require_cuda()
        Xd, arrival = _upload_blocks(X_host, n_features)
        self = cls.from_device(Xd, n_features, xyz, group=group)
        self._eng._arrival = arrival
        return self

    def _engine(self):
        if self._eng is None:
            _eng.require_cuda()
            Xd, arrival = _upload_blocks(self.X, self.n_features)
            self._eng = _eng.Engine(Xd, self.n_features, group=False)
            self._eng._arrival = arrival
        return self._eng

    def _n_rows(self):
        if self.X is not None:
            return self.X.shape[0]
        return self._eng.F * self._eng.n_c          # global rows (== local rows on a single rank)

    # ------------------------------------------------------------------ lazy host mirrors
    @property
    def X_cnt(self):
        if "X_cnt" not in self._host:
            self._host["X_cnt"] = self._eng.cnt.cpu().numpy()[:, np.newaxis]
        return self._host["X_cnt"]

    @X_cnt.setter
    def X_cnt(self, v):
        self._host["X_cnt"] = v

    @property
    def X_scl(self):
        if "X_scl" not in self._host:
            scl = self._eng.scl.cpu().numpy()
            self._host["X_scl"] = np.repeat(scl, self._eng.n_c_loc)[:, np.newaxis]
        return self._host["X_scl"]

    @X_scl.setter
    def X_scl(self, v):
        self._host["X_scl"] = v

    @property
    def X0(self):
        if "X0" not in self._host:
            self._host["X0"] = self._eng.scaled_matrix().cpu().numpy()
        return self._host["X0"]

    @X0.setter
    def X0(self, v):
        self._host["X0"] = v

    @property
    def Ur(self):
        if "Ur" not in self._host:
            self._host["Ur"] = self._eng.basis_rows().cpu().numpy()
        return self._host["Ur"]

    @Ur.setter
    def Ur(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        eng = self._engine()
        eng.set_basis_rows(torch.from_numpy(v).to(eng.dev))
        self._host["Ur"] = v

    # ------------------------------------------------------------------ scaling (a2, a12)
    def scale_data(self, scale_type='std', axis_cnt=1):
        """Centre and scale the snapshot matrix; returns X0 (sparse_sensing.py:83-171)."""
        self._scale_stats(scale_type, axis_cnt)
        return self.X0

    def _scale_stats(self, scale_type, axis_cnt, defer_row_means=False):
        if scale_type in _BROKEN_SCALES or axis_cnt not in (1, None):
            raise ValueError('could not broadcast the centring/scaling coefficient '
                             '(same failure as the reference for this option)')
        eng = self._engine()
        if scale_type == 'median':
            if eng.world > 1:        # np.median over the WHOLE feature block (:140): a local median per rank is wrong
                raise NotImplementedError("scale_type='median' needs the whole feature block on one rank")
            eng.stats('none', axis_cnt, defer_row_means)
            blocks = eng.X.view(eng.F, -1)
            k = blocks.shape[1]
            lo = torch.kthvalue(blocks, (k + 1) // 2, dim=1).values
            hi = torch.kthvalue(blocks, k // 2 + 1, dim=1).values
            eng.scl = ((lo + hi) / 2).contiguous()       # np.median (:140)
        else:
            eng.stats(scale_type, axis_cnt, defer_row_means)
        for k in ("X_cnt", "X_scl", "X0"):
            self._host.pop(k, None)

    def _sampled(self, sampling):
        """(S Ur, S X_scl, S X_cnt) on the device for a sampling matrix S (s_out x n): a row gather
        when S is one-hot (SensorMatrix or dense), a dense product otherwise (:233, :366)."""
        eng = self._engine()
        if sampling.shape[1] != self._n_rows():
            raise ValueError('The number of columns of the sampling matrix does not match the number'
                             ' of rows of X.')
        piv = _as_pivots(sampling)
        if piv is not None:
            _check_rows(piv, self._n_rows())
            if eng.world == 1:
                pd = torch.from_numpy(piv).to(eng.dev)
                SU, cnt_s = (eng.gather(pd) if eng.Ut is not None else (None, eng.cnt[pd]))
                scl_s = eng.scl[pd // eng.n_c_loc]
                return SU, scl_s, cnt_s
            sampling = _one_hot_csr(piv, self._n_rows())      # row-sharded: the same CSR path, one term per row
        # general matrix (dense, or scipy sparse such as utils.camera.project's line-of-sight CSR): never densified
        SU, cnt_s, scl_s = eng.csr_apply(sampling)
        return SU, scl_s, cnt_s

    def unscale_data(self, x0, sampling=None):
        """x = X_scl * x0 + X_cnt, or (S X_scl) * x0 + S X_cnt (sparse_sensing.py:212-240)."""
        eng = self._engine()
        xd = torch.from_numpy(np.ascontiguousarray(x0, dtype=np.float64)).to(eng.dev)
        if sampling is not None:
            _, scl_s, cnt_s = self._sampled(sampling)
            return (scl_s * xd + cnt_s).cpu().numpy()
        return eng.unscale(xd).cpu().numpy()

    # ------------------------------------------------------------------ POD (a3, a4)
    def _choose_rank(self, exp_variance, m, select_modes, n_modes):
        if select_modes == 'variance':                    # :314-324
            if not 0 <= n_modes <= 100:
                raise ValueError('The parameter n_modes is outside the[0-100] range.')
            if n_modes == 100:
                return m
            r = 1
            while exp_variance[r - 1] < n_modes:
                r += 1
            return r
        if select_modes == 'number':                      # :326-331
            if not type(n_modes) is int:
                raise TypeError('The parameter n_modes is not an integer.')
            if not 1 <= n_modes <= m:
                raise ValueError('The parameter n_modes is outside the [1-m] range.')
            return n_modes
        raise ValueError('The select_mode value is wrong.')

    def _validate_modes(self, select_modes, n_modes):
        """Raise the reference's argument errors before any device work."""
        if select_modes == 'variance':
            if not 0 <= n_modes <= 100:
                raise ValueError('The parameter n_modes is outside the[0-100] range.')
        elif select_modes == 'number':
            if not type(n_modes) is int:
                raise TypeError('The parameter n_modes is not an integer.')
            if not 1 <= n_modes <= self._m():
                raise ValueError('The parameter n_modes is outside the [1-m] range.')
        else:
            raise ValueError('The select_mode value is wrong.')

    def _m(self):
        return self.X.shape[1] if self.X is not None else self._eng.m

    def _pod(self, eng, select_modes, n_modes, centred, scaled):
        G = eng.gram(centred=centred, scaled=scaled)
        S, V = eng.eig_pod(G)
        S_h = None
        if select_modes == 'number':                          # r known: no host round trip yet
            r = self._choose_rank(None, eng.m, select_modes, n_modes)
        else:
            S_h = S.cpu().numpy()
            lam = S_h ** 2
            r = self._choose_rank(100 * np.cumsum(lam) / np.sum(lam), eng.m, select_modes, n_modes)
        # modes the Gram route cannot resolve (sigma <= sqrt(m eps) sigma_1: their eigenvalue is rounding noise of
        # G) get a zero weight in eig_pod; row-centred data always has one (rank m - 1)
        eng.backproject(eng.pod_weights[:, :r].contiguous(), centred=centred, scaled=scaled)
        sv_h = eng.pod_sv.cpu().numpy()                       # sigma | V in one D2H
        eng.check_p2p()
        m = eng.m
        S_h = sv_h[:m].copy()
        V_h = sv_h[m:].reshape(m, m).copy()
        Vr_h = V_h[:, :r].copy()
        floor = float(np.sqrt(m * _eng.EPS))
        resolved = S_h > S_h[0] * floor
        unresolved = int(np.count_nonzero(~resolved[:r]))
        if centred and r == m and not resolved[m - 1]:
            unresolved -= 1                                   # the structural null mode of row-centred data
        live = S_h[:r][resolved[:r]]
        s_min = float(live[-1]) if live.size else float(S_h[0])
        # Gram route: sigma_r to eps (sigma_1/sigma_r)^2.  Correction on the r retained modes: eps sigma_1/sigma_r,
        # valid while the subspace itself is good, (eps (sigma_1/sigma_r)^2)^2 small.  Beyond that the Gram of X0
        # does not even contain the small modes: full-width route.
        b0 = float(_eng.EPS * (S_h[0] / s_min) ** 2) if s_min > 0 else 0.0
        bound = b0
        self.pod_refined = False
        want = self.pod_refine
        if want == 'full' or (want in ('auto', True) and (
                unresolved > 0 or (b0 > self.pod_refine_tol and max(b0 * b0, _eng.EPS * S_h[0] / s_min) > 1e-10))):
            S_h, Vr_h, bound = self._pod_full_width(eng, S_h, V_h, r, centred, scaled)
            self.pod_refined = 'full'
        elif want is True or (want == 'auto' and b0 > self.pod_refine_tol):
            # One CholeskyQR2-style correction on the basis itself:  U1 = X0 V_r S^-1 = P S1 Q^T
            # (from H = U1^T U1 = Q S1^2 Q^T);  X0 V_r = P (S1 Q^T S) = P B;  SVD B = A1 S2 A2^T  =>
            # U = U1 (Q S1^-1 A1),  sigma = S2,  V = V_r A2.  Two extra passes over the n x r basis.
            H, U1 = eng.basis_gram()
            lam1, Q1 = np.linalg.eigh(H.cpu().numpy())
            S1 = np.sqrt(np.maximum(lam1, 0.0))
            ok = S1 > 0
            B = (S1[:, None] * Q1.T) * S_h[:r][None, :]
            A1, S2, A2t = np.linalg.svd(B)
            M = (Q1 * np.where(ok, 1.0 / np.where(ok, S1, 1.0), 0.0)[None, :]) @ A1
            Vn = Vr_h @ A2t.T
            sgn = np.sign(Vn[np.argmax(np.abs(Vn), axis=0), np.arange(r)])
            sgn[sgn == 0] = 1.0
            Vr_h, M = Vn * sgn, M * sgn
            eng.basis_rotate(U1, torch.from_numpy(np.ascontiguousarray(M)).to(eng.dev))
            S_h[:r] = S2
            live = S2[S2 > S2[0] * floor]
            bound = float(max(_eng.EPS * S2[0] / live[-1], b0 * b0)) if live.size else 0.0
            self.pod_refined = True
        lam = S_h ** 2
        exp_variance = 100 * np.cumsum(lam) / np.sum(lam)     # :274-275
        Ar = Vr_h * S_h[:r]                                   # A = V Sigma (:273)
        self.r = r
        self._host.pop("Ur", None)
        self.pod_sigma = S_h
        self.pod_rel_err_bound = bound
        if bound > 1e-10:
            import warnings
            warnings.warn("POD: estimated relative error of the smallest retained singular value is %.1e "
                          "(sigma_r/sigma_1 = %.1e)" % (bound, S_h[r - 1] / S_h[0]), RuntimeWarning)
        elif self.pod_refined == 'full' and S_h[r - 1] < 1e-6 * S_h[0]:
            # not a defect of this path: ANY backward-stable SVD (LAPACK's included) carries eps sigma_1/sigma_r
            self.pod_rel_err_vs_backward_stable = float(50 * _eng.EPS * S_h[0] / max(S_h[r - 1], 1e-300))
        return Ar, exp_variance[:r]

    def _pod_full_width(self, eng, S_h, V_h, r, centred, scaled):
        """POD of a matrix whose retained singular values reach below what a Gram matrix can hold
        (sigma_r/sigma_1 < ~1e-5: lambda_r/lambda_1 is at or under the rounding level of G).  The accuracy
        class of a Householder/TSQR route from Gram-kernel passes only (shifted-CholeskyQR3 style, with
        eigen-decompositions in place of Cholesky factors because row-centred X0 is exactly rank deficient):

            X0 = Y C,   Y_0 = X0 V S_f^-1,  C_0 = S_f V^T       (S_f = sigma floored at sqrt(m eps) sigma_1: exact
                                                                 identity for ANY orthogonal V, however inaccurate)
            H = Y^T Y = Q S1^2 Q^T  ->  Y <- Y Q S1_f^-1,  C <- S1_f Q^T C       (cond(Y) ~ 1e7 -> ~1e2 -> ~1)
            cond(Y) small:  SVD C = A1 S2 A2^T  =>  U = Y A1[:, :r],  sigma = S2,  V = A2

        All m columns are carried (n x m basis passes, two or three of them) -- only this hard case pays.
        Returns (sigma (m,), V_r (m, r), bound)."""
        m = eng.m
        floor = float(np.sqrt(m * _eng.EPS))
        Sf = np.maximum(S_h, floor * S_h[0])
        Cm = Sf[:, None] * V_h.T
        dev = eng.dev
        eng.backproject(torch.from_numpy(np.ascontiguousarray(V_h / Sf[None, :])).to(dev), centred=centred, scaled=scaled,
                        norms=False)
        cond = np.inf
        for it in range(4):
            H, Y = eng.basis_gram()
            lam, Q = np.linalg.eigh(H.cpu().numpy())
            lam, Q = lam[::-1], Q[:, ::-1]
            S1 = np.sqrt(np.maximum(lam, 0.0))
            S1f = np.maximum(S1, floor * S1[0])
            res = S1 > floor * S1[0]
            if centred and not res[-1]:
                res[-1] = True                                # the structural null direction stays tiny: not a defect
                cond_vals = S1[:-1][res[:-1]]
            else:
                cond_vals = S1[res]
            unresolved = int(np.count_nonzero(~res))
            cond = float(cond_vals[0] / cond_vals[-1]) if cond_vals.size else 1.0
            B = (S1f[:, None] * Q.T) @ Cm
            T = Q / S1f[None, :]
            if (unresolved == 0 and cond * cond * _eng.EPS < 1e-13) or it == 3:
                A1, S2, A2t = np.linalg.svd(B)
                Vn = A2t.T[:, :r]
                sgn = np.sign(Vn[np.argmax(np.abs(Vn), axis=0), np.arange(r)])
                sgn[sgn == 0] = 1.0
                M = (T @ A1[:, :r]) * sgn
                eng.basis_rotate(Y, torch.from_numpy(np.ascontiguousarray(M)).to(dev))
                self.pod_passes = it + 1
                bound = float(max(cond * cond * _eng.EPS, m * _eng.EPS))
                if unresolved:
                    bound = 1.0
                return S2.copy(), Vn * sgn, bound
            eng.basis_rotate(Y, torch.from_numpy(np.ascontiguousarray(T)).to(dev), norms=False)
            Cm = B
            del Y

    def decomposition(self, X0, select_modes='variance', n_modes=99):
        """POD of a scaled matrix X0 (n, m): returns (Ur, Ar, exp_variance[:r])
        (sparse_sensing.py:242-279)."""
        _eng.require_cuda()
        self._validate_modes(select_modes, n_modes)
        dev = torch.device("cuda", torch.cuda.current_device())
        X0d = torch.from_numpy(np.ascontiguousarray(X0, dtype=np.float64)).to(dev)
        tmp = _eng.Engine(X0d, 1, group=False)
        Ar, ev = self._pod(tmp, select_modes, n_modes, centred=False, scaled=False)
        Ur = tmp.basis_rows().cpu().numpy()
        return Ur, Ar, ev

    def reduction(self, U, A, exp_variance, select_modes, n_modes):
        """Truncate a basis to r modes (sparse_sensing.py:281-340)."""
        r = self._choose_rank(exp_variance, A.shape[1], select_modes, n_modes)
        if select_modes == 'number' and not 1 <= n_modes <= U.shape[1]:
            raise ValueError('The parameter n_modes is outside the [1-m] range.')
        self.r = r
        return U[:, :r], A[:, :r]

    # ------------------------------------------------------------------ fit (a5)
    def fit(self, scale_type='std', axis_cnt=1, select_modes='variance', n_modes=99, basis=None):
        """Scale, decompose (or take `basis=(Ur, Ar)`) and truncate (sparse_sensing.py:463-511)."""
        if basis is None:
            self._validate_modes(select_modes, n_modes)
        self.scale_type = scale_type
        self._scale_stats(scale_type, axis_cnt, defer_row_means=basis is None)
        eng = self._engine()
        if basis is None:
            Ar, _ = self._pod(eng, select_modes, n_modes, centred=True, scaled=True)
        else:
            self.Ur = basis[0]
            Ar = np.asarray(basis[1])
        self.Ar = Ar
        self.r = Ar.shape[1]
        sig = np.linalg.norm(Ar, axis=0)                  # :504-507 (m x r, host-trivial)
        self.Sigma_r = sig
        self.Vr = Ar / np.where(sig > 0, sig, 1.0)        # an exactly-zero mode (dropped null direction) stays zero

    # ------------------------------------------------------------------ reconstruct (a11)
    def reconstruct(self, Ar, sampling=None, *, out=None, chunk_rows=None):
        """X_rec (n, N) = unscale(Ur @ Ar.T) (sparse_sensing.py:342-375).

        The n x N result never exists on the device: row chunks of it are produced by the GEMM kernel
        into two device buffers and leave through two pinned host buffers while the next chunk is
        computed (config 4: 16.2M rows x 65 536 vectors = 8.5 TB).  Extensions (keyword-only):
          out         None: a new (n, N) numpy array is returned, like the reference;
                      an ndarray / np.memmap of shape (n, N): filled in place and returned;
                      a callable out(row0, block): called once per chunk with a (rows, N) view that is
                      only valid during the call (streams to disk, a socket, a reduction ...); returns None
          chunk_rows  rows per chunk (multiple of 128; default: ~256 MB of output per chunk)"""
        Ar = np.asarray(Ar, dtype=np.float64)
        if Ar.ndim < 2:
            Ar = Ar[np.newaxis, :]
        eng = self._engine()
        Ad = torch.from_numpy(np.ascontiguousarray(Ar)).to(eng.dev)
        if sampling is not None:                          # :365-368: (S Ur) Ar^T, sampled unscaling
            SU, scl_s, cnt_s = self._sampled(sampling)
            return (scl_s[:, None] * (SU @ Ad.T) + cnt_s[:, None]).cpu().numpy()
        n, N = eng.n_loc, int(Ad.shape[0])
        sink = out if callable(out) else None
        if sink is None:
            if out is None:
                out = np.empty((n, N))
            elif out.shape != (n, N) or out.dtype != np.float64:
                raise ValueError('out must be a float64 array of shape (n, N)')
        for row0, block in self.reconstruct_chunks(Ad, chunk_rows=chunk_rows):
            if sink is not None:
                sink(row0, block)
            else:
                out[row0:row0 + block.shape[0]] = block
        return None if sink is not None else out

    def reconstruct_chunks(self, Ar, chunk_rows=None):
        """Generator over (row0, block): block = rows [row0, row0 + rows) of the reconstruction as a
        (rows, N) numpy view of a pinned host buffer, valid until the next iteration.  Two device
        buffers and two pinned buffers: chunk k+1 is computed while chunk k crosses PCIe."""
        eng = self._engine()
        if torch.is_tensor(Ar):
            Ad = Ar.to(eng.dev, torch.float64)
        else:
            Ar = np.asarray(Ar, dtype=np.float64)
            Ad = torch.from_numpy(np.ascontiguousarray(Ar[np.newaxis, :] if Ar.ndim < 2 else Ar)).to(eng.dev)
        if Ad.shape[1] != eng.r:
            raise ValueError('The number of columns of Ar does not match the number of modes.')
        n, N = eng.n_loc, int(Ad.shape[0])
        if chunk_rows is None:
            chunk_rows = max(128, ((256 << 20) // (8 * N)) // 128 * 128)
        chunk_rows = int(min(max(128, (int(chunk_rows) + 127) // 128 * 128), (n + 127) // 128 * 128))
        nbuf = 1 if chunk_rows >= n else 2
        key = (chunk_rows, N, nbuf)
        st = getattr(self, "_recon_ring", None)
        if st is None or st["key"] != key:
            st = {"key": key,
                  "dev": [torch.empty(chunk_rows, N, dtype=torch.float64, device=eng.dev) for _ in range(nbuf)],
                  "host": [torch.empty(chunk_rows, N, dtype=torch.float64, pin_memory=chunk_rows * N >= (1 << 17))
                           for _ in range(nbuf)],
                  "copy": torch.cuda.Stream(device=eng.dev)}
            self._recon_ring = st
        cur = torch.cuda.current_stream(eng.dev)
        copy = st["copy"]
        done = [None] * nbuf          # D2H of the buffer's previous chunk finished
        pending = None                # (row0, rows, buffer) of the chunk in flight to the host
        for k, row0 in enumerate(range(0, n, chunk_rows)):
            b = k % nbuf
            rows = min(chunk_rows, n - row0)
            if done[b] is not None:
                cur.wait_event(done[b])                   # the device buffer is free again
            eng.reconstruct(Ad, row0=row0, nrows=rows, out=st["dev"][b])
            ready = torch.cuda.Event()
            ready.record(cur)
            if pending is not None:                       # hand the previous chunk to the caller while
                p0, prow, pb = pending                    # this one is being computed
                done[pb].synchronize()
                yield p0, st["host"][pb][:prow].numpy()
            copy.wait_event(ready)
            with torch.cuda.stream(copy):
                st["host"][b][:rows].copy_(st["dev"][b][:rows], non_blocking=True)
                done[b] = torch.cuda.Event()
                done[b].record(copy)
            pending = (row0, rows, b)
        if pending is not None:
            p0, prow, pb = pending
            done[pb].synchronize()
            yield p0, st["host"][pb][:prow].numpy()

    # ------------------------------------------------------------------ out of scope
    def scale_limits(self, limits):
        raise NotImplementedError('scale_limits (COLS only) is out of scope, SURVEY.md 8')

    def adaptive_sampling(self, P, scale_type='std'):
        raise NotImplementedError('adaptive_sampling is out of scope, SURVEY.md 8')

    def CPOD(self, *args, **kwargs):
        raise NotImplementedError('CPOD (cvxpy) is out of scope, SURVEY.md 8')


class SPR(ROM):
    """Sparse Placement for Reconstruction (reference SPR, sparse_sensing.py:513-901)."""

    def __init__(self, X, n_features, xyz):
        super().__init__(X, n_features, xyz)

    # ------------------------------------------------------------------ GEM placement (8f row 1)
    def gem(self, Ur, n_sensors, mask, d_min, verbose):
        """Greedy entropy-maximisation placement (sparse_sensing.py:586-698): the row indices of the
        n_sensors sensors.  Ur: (n, r) basis (None: the fitted basis already on the device).  The
        reference's random covariance jitter (:667) is drawn with the same np.random.normal calls, so
        seeding numpy's global generator reproduces the reference's selection."""
        eng = self._engine()
        Ut = None
        if Ur is not None and Ur is not self._host.get("Ur"):
            Urd = torch.from_numpy(np.ascontiguousarray(Ur, dtype=np.float64)).to(eng.dev)
            if Urd.shape[0] != eng.n_loc:
                raise ValueError('Ur must have one row per row of X')
            Ut = eng.tiled_copy(Urd)
        elif eng.Ut is None:
            raise ValueError('gem() needs a fitted basis')
        mask_d = None if mask is None else torch.from_numpy(np.asarray(mask, dtype=bool)).to(eng.dev)
        xyz_d = None
        if d_min > 0:
            if self.xyz is None:
                raise TypeError('d_min > 0 needs the cell coordinates xyz')
            xyz_d = torch.from_numpy(np.ascontiguousarray(self.xyz, dtype=np.float64)).to(eng.dev)
            if xyz_d.shape != (eng.n_c_loc, 3):
                raise ValueError('xyz must be (n_points, 3)')
        return eng.gem(n_sensors, mask_d, xyz_d, float(d_min), Ut=Ut, verbose=verbose)

    # ------------------------------------------------------------------ placement (a7)
    def optimal_placement(self, calc_type='qr', n_sensors=10, mask=None, d_min=0., verbose=False,
                          block=8):
        """QR-with-column-pivoting sensor placement (sparse_sensing.py:700-756, 'qr' branch).
        `n_sensors`, `d_min`, `verbose` are ignored for 'qr' exactly as in the reference.
        `block` (extension) = pivot steps between trailing-matrix rewrites."""
        if calc_type == 'gem':                            # :745-751
            P = self.gem(self.Ur if "Ur" in self._host else None, n_sensors, mask, d_min, verbose)
            return SensorMatrix(P, self._n_rows())
        if calc_type != 'qr':
            raise NotImplementedError('The sensor selection method has not been implemented yet')
        eng = self._engine()
        if mask is not None:                              # :737-738, mutates the basis
            eng.mask_rows(torch.from_numpy(np.asarray(mask, dtype=bool)).to(eng.dev))
            self._host.pop("Ur", None)
        piv, rdiag, gap = eng.qrcp(block=block)
        out = eng.qr_out.cpu()                            # pivots | R diagonal | gaps in one D2H
        s_ = int(piv.numel())
        self.qr_pivots = out[:s_].view(torch.int64).numpy().copy()
        eng.check_p2p()
        self.qr_rdiag = out[s_:2 * s_].numpy().copy()
        self.qr_gap = out[2 * s_:].numpy().copy()
        return SensorMatrix(self.qr_pivots, self._n_rows())

    # ------------------------------------------------------------------ train (a8)
    def train(self, C, is_Theta=False, limits=None, method='OLS', solver='CLARABEL', cond=False,
              verbose=False):
        """Theta = C . Ur (sparse_sensing.py:758-820)."""
        n = self._n_rows()
        if (C.shape[1] != n) and not is_Theta:
            raise ValueError('The number of columns of C does not match the number'
                             ' of rows of X.')
        eng = self._engine()
        if not is_Theta:
            self.C = C
            piv = _as_pivots(C)
            if piv is not None:
                _check_rows(piv, n)
                Theta_d, cnt_s = eng.gather(torch.from_numpy(piv).to(eng.dev))
            else:                                          # general C: CSR kernel, row-sharded like everything else
                if eng.Ut is None:
                    raise AttributeError('The function fit has to be called before calling train.')
                Theta_d, cnt_s, _ = eng.csr_apply(C)
            self._cnt_s = cnt_s
        else:
            Theta_d = torch.from_numpy(np.ascontiguousarray(C, dtype=np.float64)).to(eng.dev)
        if Theta_d.shape[1] != eng.r:
            raise ValueError('The number of columns of Theta does not match the number'
                             ' of columns of Ur.')
        self._Theta_d = Theta_d
        with _eng.LINALG_LOCK:
            self._PinvT = torch.linalg.pinv(Theta_d, rtol=1e-15).T.contiguous()
        self.Theta = Theta_d.cpu().numpy()
        self.limits = limits
        self.method = method
        self.solver = solver
        self.verbose = verbose
        if cond == True:                                  # :813-820
            with _eng.LINALG_LOCK:
                Sth = torch.linalg.svdvals(Theta_d if Theta_d.shape[0] == Theta_d.shape[1]
                                           else torch.linalg.pinv(Theta_d, rtol=1e-15))
            self.k = float(Sth[0] / Sth[-1])

    # ------------------------------------------------------------------ predict (a9, a10)
    def scale_vector(self, y):
        """y0[:,0] = (y[:,0] - C X_cnt)/scl, y0[:,1] = y[:,1]/scl (sparse_sensing.py:553-584)."""
        eng = self._engine()
        y0 = np.zeros((y.shape[0], 2))
        cnt_vector = self._cnt_s.cpu().numpy()
        scl_h = eng.scl.cpu().numpy()
        scl_vector = scl_h[y[:, 2].astype('int')]          # X_scl[feature * n_points, 0] (:576)
        y0[:, 0] = (y[:, 0] - cnt_vector) / scl_vector
        y0[:, 1] = y[:, 1] / scl_vector
        self.cnt_vector = cnt_vector
        self.scl_vector = scl_vector
        return y0

    def predict(self, y):
        """OLS estimate of the POD coefficients from sparse measurements
        (sparse_sensing.py:822-901).  Returns (Ar (N, r), Ar_sigma (N, r))."""
        if isinstance(y, np.ndarray):
            y = [y]
        if not hasattr(self, 'Theta'):
            raise AttributeError('The function fit has to be called '
                                 'before calling predict.')
        for yi in y:
            if self.Theta.shape[0] != yi.shape[0]:
                raise ValueError('The number of rows of Theta does not match the number'
                                 ' of rows of y.')
            if yi.shape[1] != 3:
                raise ValueError('The y array has the wrong number of columns. y has'
                                 ' to have dimensions (s,3).')
        if self.method == 'COLS':
            raise NotImplementedError('constrained OLS (cvxpy) is out of scope, SURVEY.md 8')
        if self.method != 'OLS':
            raise NotImplementedError('The prediction method selected has not been '
                                      'implemented yet')
        eng = self._engine()
        N = len(y)
        Y = np.stack([np.asarray(yi, dtype=np.float64) for yi in y])        # (N, s, 3)
        fid = Y[:, :, 2]
        if fid.size and (fid.min() < 0 or fid.max() >= self.n_features):     # the reference's X_scl[id * n_points] (:576)
            raise IndexError('feature index out of range in y[:, 2]')
        Yd = torch.from_numpy(Y).to(eng.dev)
        scl_s = eng.scl[Yd[:, :, 2].to(torch.int64)]                         # (N, s)
        cnt_s = self._cnt_s
        weighted = torch.any(Yd[:, :, 1] != 0, dim=1)                        # :868
        Ar = torch.zeros(N, eng.r, dtype=torch.float64, device=eng.dev)
        Asig = torch.zeros(N, eng.r, dtype=torch.float64, device=eng.dev)
        plain = ~weighted
        if bool(plain.any()):
            idx = torch.nonzero(plain).flatten()
            same_features = bool((Yd[idx, :, 2] == Yd[idx[0], :, 2]).all())
            if same_features:
                Ar[idx] = eng.ols_predict(Yd[idx, :, 0].contiguous(), cnt_s,
                                          scl_s[idx[0]].contiguous(), self._PinvT)
            else:
                Y0 = (Yd[idx, :, 0] - cnt_s) / scl_s[idx]
                Ar[idx] = eng.ols_predict(Y0.contiguous(), None, None, self._PinvT)
        if bool(weighted.any()):                                             # :871-878
            idx = torch.nonzero(weighted).flatten()
            y0v = ((Yd[idx, :, 0] - cnt_s) / scl_s[idx]).contiguous()
            y0s = (Yd[idx, :, 1] / scl_s[idx]).contiguous()
            a_w, s_w = eng.wols_predict(self._Theta_d, y0v, y0s)
            Ar[idx] = a_w
            Asig[idx] = s_w
        self.scale_vector(y[-1])                # leaves cnt_vector / scl_vector like the reference
        return Ar.cpu().numpy(), Asig.cpu().numpy()


# ------------------------------------------------------------------------------------------------------
# one process, several GPUs, the unchanged constructor (ROM.devices / OMB_DEVICES; see multi.py)
# ------------------------------------------------------------------------------------------------------
_REPLICATED = ("scale_type", "Ar", "r", "Sigma_r", "Vr", "pod_sigma", "pod_rel_err_bound", "pod_refined", "qr_pivots",
               "qr_rdiag", "qr_gap", "Theta", "k", "limits", "method", "solver", "verbose", "cnt_vector", "scl_vector")


class _MultiROM(ROM):
    _single = ROM

    def __init__(self, X, n_features, xyz):
        ROM.__init__(self, X, n_features, xyz)
        from . import multi as _multi
        _eng.require_cuda()
        self._md = _multi.MultiDevice(self._single, X, n_features, xyz, _multi.resolve_devices(type(self).devices))
        for s in self._md.subs:
            s.pod_refine, s.pod_refine_tol = self.pod_refine, self.pod_refine_tol

    def _engine(self):
        raise RuntimeError('internal: the multi-device object has one engine per device')

    def _pull(self):
        s0 = self._md.subs[0]
        for k in _REPLICATED:
            if hasattr(s0, k):
                setattr(self, k, getattr(s0, k))

    def _rows(self, name):
        return self._md.assemble(self._md.run(lambda s, g: getattr(s, name)))

    X_cnt = property(lambda self: self._rows("X_cnt"))
    X_scl = property(lambda self: self._rows("X_scl"))
    X0 = property(lambda self: self._rows("X0"))

    @property
    def Ur(self):
        return self._rows("Ur")

    @Ur.setter
    def Ur(self, v):
        v = np.asarray(v, dtype=np.float64)
        self._md.run(lambda s, g: setattr(s, "Ur", self._md.shard(v, g)))

    def scale_data(self, scale_type='std', axis_cnt=1):
        self._md.call("_scale_stats", scale_type, axis_cnt)
        return self.X0

    def unscale_data(self, x0, sampling=None):
        if sampling is not None:
            return self._md.call("unscale_data", x0, sampling)[0]
        x0 = np.asarray(x0, dtype=np.float64)
        return self._md.assemble(self._md.run(lambda s, g: s.unscale_data(self._md.shard(x0, g))))

    def decomposition(self, X0, select_modes='variance', n_modes=99):
        out = self._md.run(lambda s, g: s.decomposition(X0, select_modes, n_modes) if g == 0 else None)[0]
        self.r = self._md.subs[0].r
        return out

    def fit(self, scale_type='std', axis_cnt=1, select_modes='variance', n_modes=99, basis=None):
        if basis is None:
            self._validate_modes(select_modes, n_modes)
        for s in self._md.subs:
            s.pod_refine, s.pod_refine_tol = self.pod_refine, self.pod_refine_tol
        U = None if basis is None else np.asarray(basis[0], dtype=np.float64)
        self._md.run(lambda s, g: s.fit(scale_type, axis_cnt, select_modes, n_modes,
                                        None if basis is None else (self._md.shard(U, g), basis[1])))
        self._pull()

    def _pieces(self, g, row0, rows):
        """Local rows [row0, row0 + rows) of device g as (local offset, count, global row0) runs."""
        ncl, c0, n_c = self._md.cells[g], self._md.offsets[g], self._md.n_c
        out, lo = [], row0
        while lo < row0 + rows:
            f = lo // ncl
            hi = min(row0 + rows, (f + 1) * ncl)
            out.append((lo - row0, hi - lo, f * n_c + c0 + (lo - f * ncl)))
            lo = hi
        return out

    def reconstruct(self, Ar, sampling=None, *, out=None, chunk_rows=None):
        if sampling is not None:
            return self._md.call("reconstruct", Ar, sampling)[0]
        Ar = np.asarray(Ar, dtype=np.float64)
        Ar = Ar[np.newaxis, :] if Ar.ndim < 2 else Ar
        n, N = self._md.F * self._md.n_c, Ar.shape[0]
        sink = out if callable(out) else None
        if sink is None:
            if out is None:
                out = np.empty((n, N))
            elif out.shape != (n, N) or out.dtype != np.float64:
                raise ValueError('out must be a float64 array of shape (n, N)')
        import threading
        lock = threading.Lock()

        def work(s, g):
            for row0, block in s.reconstruct_chunks(Ar, chunk_rows=chunk_rows):
                for off, cnt, g0 in self._pieces(g, row0, block.shape[0]):
                    if sink is None:
                        out[g0:g0 + cnt] = block[off:off + cnt]
                    else:
                        with lock:
                            sink(g0, block[off:off + cnt])

        self._md.run(work)
        return None if sink is not None else out

    def reconstruct_chunks(self, Ar, chunk_rows=None):
        for g, s in enumerate(self._md.subs):
            with torch.cuda.device(self._md.devices[g]):
                for row0, block in s.reconstruct_chunks(Ar, chunk_rows=chunk_rows):
                    for off, cnt, g0 in self._pieces(g, row0, block.shape[0]):
                        yield g0, block[off:off + cnt]


class _MultiSPR(_MultiROM, SPR):
    _single = SPR

    def __init__(self, X, n_features, xyz):
        _MultiROM.__init__(self, X, n_features, xyz)

    def gem(self, Ur, n_sensors, mask, d_min, verbose):
        U = None if Ur is None else np.asarray(Ur, dtype=np.float64)
        mk = None if mask is None else np.asarray(mask, dtype=bool)
        return self._md.run(lambda s, g: s.gem(None if U is None else self._md.shard(U, g), n_sensors,
                                               None if mk is None else self._md.shard(mk, g), d_min,
                                               verbose and g == 0))[0]

    def optimal_placement(self, calc_type='qr', n_sensors=10, mask=None, d_min=0., verbose=False, block=8):
        if calc_type not in ('qr', 'gem'):
            raise NotImplementedError('The sensor selection method has not been implemented yet')
        mk = None if mask is None else np.asarray(mask, dtype=bool)
        C = self._md.run(lambda s, g: s.optimal_placement(calc_type, n_sensors, None if mk is None else self._md.shard(mk, g),
                                                          d_min, verbose and g == 0, block))[0]
        self._pull()
        return C

    def train(self, C, is_Theta=False, limits=None, method='OLS', solver='CLARABEL', cond=False, verbose=False):
        n = self._md.F * self._md.n_c
        if (C.shape[1] != n) and not is_Theta:
            raise ValueError('The number of columns of C does not match the number'
                             ' of rows of X.')
        self._md.call("train", C, is_Theta, limits, method, solver, cond, verbose)
        if not is_Theta:
            self.C = C
        self._pull()

    def scale_vector(self, y):
        out = self._md.run(lambda s, g: s.scale_vector(y) if g == 0 else None)[0]
        self._pull()
        return out

    def predict(self, y):
        if not hasattr(self, 'Theta'):
            raise AttributeError('The function fit has to be called '
                                 'before calling predict.')
        out = self._md.run(lambda s, g: s.predict(y) if g == 0 else None)[0]      # Theta is replicated: no exchange
        self._pull()
        return out
