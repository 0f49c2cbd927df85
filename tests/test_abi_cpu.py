"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol declared in
include/omb200.h (no compute calls), and the host mirror raises the reference's errors before any
device work (reference sparse_sensing.py:69-81, :314-333, :752-754)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from openmeasure_b200 import build
    build.build()
    from openmeasure_b200 import _lib
    return _lib


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "omb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(omb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    L = ctypes.CDLL(lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/omb200.h but not exported"


def test_bindings_cover_the_header(lib):
    assert sorted(lib.SIGNATURES) == _declared_symbols()
    L = lib.load()
    assert L.omb_version() >= 100
    assert L.omb_launch_count() == 0


def test_argument_errors_precede_device_work(lib):
    L = lib.load()
    rc = L.omb_row_means(None, 4, 4, None, None)
    assert rc < 0 and b"null pointer" in L.omb_last_error()
    rc = L.omb_qrcp(None, 10, 4, 4, None, None, None, 1, 0, None, None, None, None)
    assert rc < 0
    assert L.omb_launch_count() == 0


def test_constructor_and_mode_errors_match_reference():
    from openmeasure_b200.sparse_sensing import ROM, SPR
    X = np.zeros((6, 3))
    with pytest.raises(TypeError):
        ROM([[1.0, 2.0]], 1, None)
    with pytest.raises(TypeError):
        ROM(X, 2.0, None)
    with pytest.raises(Exception):
        SPR(X, 4, None)
    rom = ROM(X, 2, None)
    assert rom.n_points == 3 and rom.X is X
    with pytest.raises(ValueError):
        rom.fit(select_modes='variance', n_modes=101)
    with pytest.raises(TypeError):
        rom.fit(select_modes='number', n_modes=2.0)
    with pytest.raises(ValueError):
        rom.fit(select_modes='number', n_modes=4)
    with pytest.raises(ValueError):
        rom.fit(select_modes='bogus', n_modes=1)
    ev = np.array([50.0, 90.0, 100.0])
    U, A = rom.reduction(np.zeros((6, 3)), np.zeros((3, 3)), ev, 'variance', 80)
    assert rom.r == 2 and U.shape == (6, 2) and A.shape == (3, 2)


def test_sensor_matrix_supports_reference_idioms():
    from openmeasure_b200.sparse_sensing import SensorMatrix
    C = SensorMatrix([4, 0, 7], 9)
    assert C.shape == (3, 9)
    assert np.argmax(C[1, :]) == 0 and np.argmax(C[2, :]) == 7
    x = np.arange(9.0) * 2
    np.testing.assert_array_equal(C @ x, [8.0, 0.0, 14.0])
    np.testing.assert_array_equal(C.dot(x), np.asarray(C) @ x)
    dense = np.asarray(C)
    assert dense.shape == (3, 9) and dense.sum() == 3 and dense[0, 4] == 1


def test_no_cuda_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from openmeasure_b200 import _lib
    from openmeasure_b200.sparse_sensing import SPR
    spr = SPR(np.random.default_rng(0).random((20, 5)), 2, None)
    with pytest.raises(_lib.OmbError):
        spr.fit(n_modes=100)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "openmeasure_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "oracle.clib" not in src and "pod_oracle" not in src, f


# ---- ingest: .npy header parsing and shard byte ranges (host logic, no device) -----------------
def test_npy_shard_ranges(tmp_path):
    import numpy as np
    from openmeasure_b200 import ingest
    X = np.arange(3 * 10 * 4, dtype=np.float64).reshape(30, 4)
    p = tmp_path / "X.npy"
    np.save(p, X)
    shape, off = ingest.npy_header(str(p))
    assert shape == (30, 4)
    raw = open(p, "rb").read()
    got = []
    for rank in range(3):
        ranges, ncl = ingest.shard_ranges(shape, off, 3, rank, 3)
        c0, ncl2 = ingest.shard_cells(10, rank, 3)
        assert ncl == ncl2
        rows = np.concatenate([np.frombuffer(raw[o:o + nb], dtype=np.float64).reshape(-1, 4) for o, nb, _ in ranges])
        np.testing.assert_array_equal(rows, np.concatenate([X[f * 10 + c0: f * 10 + c0 + ncl] for f in range(3)]))
        got.append(ncl)
    assert sum(got) == 10 and got == [4, 3, 3]
    np.save(p, X.astype(np.float32))
    with pytest.raises(ValueError):
        ingest.npy_header(str(p))
    np.save(p, np.asfortranarray(X))
    with pytest.raises(ValueError):
        ingest.npy_header(str(p))


def test_reference_import_line_resolves_to_the_b200_classes():
    """`from openmeasure.sparse_sensing import ROM, SPR` (the reference's README import) works unchanged."""
    from openmeasure.sparse_sensing import ROM, SPR
    from openmeasure_b200 import sparse_sensing as b
    assert ROM is b.ROM and SPR is b.SPR and issubclass(SPR, ROM)
