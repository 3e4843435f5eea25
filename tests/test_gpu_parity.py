"""GPU parity tests (run on the B200 box with ``-m gpu``): the CUDA path, called through the C ABI, against the
CPU oracle on the same seeded inputs and against the golden vectors the reference itself produced
(tests/golden/*.npz).  Tolerances:

* P-f64   1e-9 relative on float64 arrays (A, mBhalf, pre-cast T, D/N/E, node weights),
* P-f32   2e-6 relative on the float32 products the reference emits (T, UC, Sigma, kappa, outimage, Tsum, Neff),
* P-discrete  identical bracket index / branch word / CG iteration counts.
"""

import os
import warnings

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from oracle import lakernel as OL  # noqa: E402
from oracle import routines as R  # noqa: E402
from oracle.sysmat import OracleOutStamp  # noqa: E402
from pyimcom_b200 import _lib  # noqa: E402
from pyimcom_b200 import lakernel as GL  # noqa: E402
from pyimcom_b200 import pyimcom_croutines as G  # noqa: E402
from pyimcom_b200.coadd import GpuBlock, GpuOutStamp  # noqa: E402
from pyimcom_b200.psfovl_host import PSFTables  # noqa: E402

P64 = 1e-9
P32 = 2e-6


def rel(a, b):
    b = np.asarray(b, dtype=np.float64)
    return np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def gr(golden_dir):
    return np.load(os.path.join(golden_dir, "routine.npz"))


# ---------------------------------------------------------------------------------------------------
# function seam (furry_parakeet.pyimcom_croutines)
# ---------------------------------------------------------------------------------------------------
def test_interp_seam(gr):
    """tests/pyimcom/test_routine.py:8-63 through the GPU library (reference tolerance there: 1e-9)."""
    infunc, x_, y_, xs_, ys_, xpos, ypos = cases.interp_inputs()
    f = np.zeros((2, x_.size))
    G.iD5512C(infunc, x_, y_, f)
    assert np.abs(f).max() > 0.98
    assert np.abs(f - gr["iD5512C"]).max() < 1e-13
    fs = np.zeros((2, x_.size))
    G.iD5512C_sym(infunc, xs_, ys_, fs)
    assert np.abs(fs - gr["iD5512C_sym"]).max() < 1e-13
    f2 = np.zeros((2, x_.size))
    G.iD5512C(infunc, xs_, ys_, f2)
    assert np.abs(fs - f2).max() < 1e-9
    g = np.zeros((xpos.shape[0], xpos.shape[1] * ypos.shape[1]))
    G.gridD5512C(infunc[0], xpos, ypos, g)
    assert np.abs(g).max() > 0.98
    assert np.abs(g - gr["gridD5512C"]).max() < 1e-13


def test_interp_offgrid_and_empty():
    """Off-grid points: scattered output untouched (routine.py:166-167), grid output zero (routine.py:307-310)."""
    infunc = np.random.default_rng(0).standard_normal((1, 30, 40))
    x = np.array([3.99, 4.0, 20.3, 34.99, 35.0, -1.0, 1e6])
    y = np.array([10.0, 10.2, 3.5, 12.0, 12.0, 10.0, 10.0])
    out = np.full((1, x.size), 7.0)
    ref = out.copy()
    G.iD5512C(infunc, x, y, out)
    R.iD5512C(infunc, x, y, ref)
    assert np.array_equal(out == 7.0, ref == 7.0)
    assert np.abs(out - ref).max() < 1e-13
    xp = np.array([[2.0, 10.5, 36.0, 20.25]])
    yp = np.array([[1.0, 15.5, 25.5]])
    g, gref = np.full((1, 12), 5.0), np.zeros((1, 12))
    G.gridD5512C(infunc[0], xp, yp, g)
    R.gridD5512C(infunc[0], xp, yp, gref)
    assert np.abs(g - gref).max() < 1e-13 and (g == 0).sum() == (gref == 0).sum()
    e = np.zeros((1, 0))
    G.iD5512C(infunc, np.zeros(0), np.zeros(0), e)  # empty input is a no-op


def test_interp_large_random():
    """Full-size scattered call shape of the reference (129^2 points on a 395^2 table, psfutil.py:1469-1477)."""
    rng = np.random.default_rng(5)
    tab = rng.standard_normal((1, 395, 395))
    n = 129 * 129
    x, y = rng.uniform(0, 395, n), rng.uniform(0, 395, n)
    out, ref = np.zeros((1, n)), np.zeros((1, n))
    G.iD5512C(tab, x, y, out)
    R.iD5512C(tab, x, y, ref)
    assert np.abs(out - ref).max() < 1e-12
    xp, yp = np.sort(rng.uniform(0, 395, (129, 38))), np.sort(rng.uniform(0, 395, (129, 38)))
    g, gref = np.zeros((129, 38 * 38)), np.zeros((129, 38 * 38))
    G.gridD5512C(tab[0], xp, yp, g)
    R.gridD5512C(tab[0], xp, yp, gref)
    assert np.abs(g - gref).max() < 1e-12


def test_getw(gr):
    """tests/pyimcom/test_psf.py:57-63."""
    w = np.zeros(10)
    for k, fh in enumerate(cases.GETW_FH):
        G.iD5512C_getw(w, fh)
        assert np.abs(w - gr["getw"][k]).max() < 1e-15
    G.iD5512C_getw(w, 0.5)
    e5 = np.zeros(10)
    e5[5] = 1.0
    assert np.abs(w - e5).max() < 1e-8


def test_lakernel1_and_lsolve(gr):
    """tests/pyimcom/test_routine.py:66-156 through the GPU library, same tolerances as the reference's C-vs-Numba."""
    A, mBhalf, Cn = cases.kernel_toy()
    lam, Q = np.linalg.eigh(A)
    mPhalf = np.ascontiguousarray(mBhalf @ Q)
    m, n = mBhalf.shape
    kappa, Sigma, UC, T = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros((m, n))
    G.lakernel1(lam, Q, mPhalf, Cn, 1e-8, 1e-16, 1e16, 53, kappa, Sigma, UC, T, 0.5)
    assert 2.5e-7 < kappa.min() and kappa.max() < 3.5e-7
    assert 0.34 < Sigma.min() and Sigma.max() < 0.38
    assert 9e-9 < UC.min() and UC.max() < 1.1e-8
    assert 0.077 < np.abs(T).max() < 0.079
    assert np.abs(kappa - gr["lk1_kappa"]).max() < 1e-12
    assert np.abs(Sigma - gr["lk1_Sigma"]).max() < 1e-7
    assert np.abs(UC - gr["lk1_UC"]).max() < 1e-14
    # T is expressed in the eigenbasis, whose gauge (signs, rotations inside degenerate subspaces) depends on the
    # host's LAPACK build: compare it with the oracle on the SAME (lam, Q), and with the reference through the
    # gauge-invariant product T @ Q^T (lakernel.py:223)
    rk, rS, rU, rT = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros((m, n))
    R.lakernel1(lam, Q, mPhalf, Cn, 1e-8, 1e-16, 1e16, 53, rk, rS, rU, rT, 0.5)
    # 53 halvings of log(factor) end below float64 resolution, where the last branches are decided by the rounding
    # of the n-term sums: the GPU's tree reduction and the serial CPU sum may then part ways at the 1e-8 level
    assert np.abs(kappa - rk).max() < 1e-7 * rk.max()
    # P-discrete at the production depth (nbis = 13, lakernel.py:174): kappa sits on the lattice
    # sqrt(kmin kmax) * factor^(+-1/2 +-1/4 ...), so equal branch words <=> equal kappa up to rounding of the products
    k13, r13 = np.zeros(m), np.zeros(m)
    S_, U_, T13, rT13 = np.zeros(m), np.zeros(m), np.zeros((m, n)), np.zeros((m, n))
    G.lakernel1(lam, Q, mPhalf, Cn, 1e-8, 1e-16, 1e16, 13, k13, S_, U_, T13, 0.5)
    R.lakernel1(lam, Q, mPhalf, Cn, 1e-8, 1e-16, 1e16, 13, r13, S_, U_, rT13, 0.5)
    assert np.abs(k13 / r13 - 1).max() < 1e-13
    assert np.abs(T13 - rT13).max() < 1e-12 * np.abs(rT13).max()  # same kappa => T = P/(lam+kappa) to rounding
    assert np.abs(T - rT).max() < 1e-7 * np.abs(rT).max()  # nbis = 53: follows the 1e-8 kappa wobble above
    assert np.abs((T @ Q.T)[::25, ::33] - gr["lk1_TQt_sub"]).max() < 1e-8
    # float32 outputs, as lakernel.py:216-218 passes them
    k32, S32, U32 = (np.zeros(m, dtype=np.float32) for _ in range(3))
    G.lakernel1(lam, Q, mPhalf, Cn, 1e-8, 1e-16, 1e16, 53, k32, S32, U32, T, 0.5)
    assert np.array_equal(k32, kappa.astype(np.float32))
    A_ = A + np.identity(n)
    x = np.zeros(n)
    G.lsolve_sps(n, A_.copy(), x, mBhalf[0].copy())
    assert np.abs(x - np.linalg.solve(A_, mBhalf[0])).max() < 1e-10
    assert np.abs(x - gr["lsolve_x"]).max() < 1e-12


def test_build_reduced_T(gr):
    Nf, Df, Ef, kap, ucmin, smax = cases.reduced_inputs()
    m = Df.size // kap.size
    ok, oS, oU, ow = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(m * kap.size)
    iv, br = np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32)
    G.build_reduced_T_wrap(Nf, Df, Ef, kap, ucmin, smax, ok, oS, oU, ow, iv, br)
    rk, rS, rU, rw = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(m * kap.size)
    riv, rbr = np.zeros(m, dtype=np.int32), np.zeros(m, dtype=np.int32)
    R.build_reduced_T_wrap(Nf, Df, Ef, kap, ucmin, smax, rk, rS, rU, rw, riv, rbr)
    assert np.array_equal(iv, riv) and np.array_equal(br, rbr)  # P-discrete
    assert rel(ok, gr["brt_kappa"]) < 1e-14
    assert rel(oS, gr["brt_Sigma"]) < 1e-10
    assert np.abs(oU - gr["brt_UC"]).max() < 1e-10
    assert rel(ow, gr["brt_w"]) < P64


# ---------------------------------------------------------------------------------------------------
# dense linear algebra building blocks
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K", [208, 210, 8, 1042])  # whole stages, a K tail, less than one stage, a long ring
def test_gemm_nt(K):
    rng = np.random.default_rng(1)
    M, N = 256, 384
    A, B, Cm = rng.standard_normal((M, K)), rng.standard_normal((N, K)), rng.standard_normal((M, N))
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    for acc, want in ((0, A @ B.T), (1, Cm + A @ B.T), (-1, Cm - A @ B.T)):
        dC = torch.from_numpy(Cm.copy()).cuda()
        _lib.dev_gemm_nt(GL.ptr(dA), K, GL.ptr(dB), K, GL.ptr(dC), N, M, N, K, acc, GL.stream_handle())
        assert rel(dC.cpu().numpy(), want) < 1e-13


@pytest.mark.parametrize("n,m", [(100, 37), (128, 128), (700, 300), (1500, 729)])
def test_chol_solve_random(n, m):
    """Factor + solve vs SciPy on a random SPD system with a wide spectrum."""
    from scipy.linalg import cho_solve, cholesky

    rng = np.random.default_rng(n)
    Gm = rng.standard_normal((n, n))
    Q, _ = np.linalg.qr(Gm)
    A = (Q * np.logspace(-6, 0, n)) @ Q.T
    A = 0.5 * (A + A.T)
    B = rng.standard_normal((m, n))
    ref = cho_solve((cholesky(A, lower=True), True), B.T).T
    ds = GL.upload_system(A, B[None], [1.0], 1)
    W = GL._padded_system(ds, [])
    X = ds.mB[0].clone()
    info, _k = GL.chol_solve_batch([W], [X])
    assert int(info.item()) == 0
    got = X[:m, :n].cpu().numpy()
    # compare through the residual-insensitive measure used by P-f64
    assert rel(got, ref) < 1e-9
    L = np.tril(W[:n, :n].cpu().numpy())
    assert rel(L @ L.T, A) < 1e-13


@pytest.mark.parametrize("case", ["identity", "diagonal_16_decades", "scaled_1e-150", "scaled_1e+150", "graded_rows",
                                  "rhs_zero_rows"])
def test_chol_solve_special_systems(case):
    """The batched Cholesky + triangular solves (long-K updates on the sliced INT8 path: n = 1300 has three super-panels)
    on systems that stress the digit-plane scales: trivial factors, diagonal entries over 16 decades, entries near the
    ends of the float64 range, symmetric row / column scaling D A D, right-hand sides with all-zero rows."""
    from scipy.linalg import cho_solve, cholesky

    rng = np.random.default_rng(42)
    n, m = 1300, 150
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    base = (Q * np.logspace(-4, 0, n)) @ Q.T
    base = 0.5 * (base + base.T)
    B = rng.standard_normal((m, n))
    if case == "identity":
        A = np.eye(n)
    elif case == "diagonal_16_decades":
        A = np.diag(np.logspace(-8, 8, n))
    elif case == "scaled_1e-150":
        A = 1e-150 * base
    elif case == "scaled_1e+150":
        A = 1e150 * base
    elif case == "graded_rows":
        d = np.logspace(-6, 6, n)[rng.permutation(n)]
        A = d[:, None] * base * d[None, :]
    else:
        A = base
        B[::7] = 0.0
    ref = cho_solve((cholesky(A, lower=True), True), B.T).T
    ds = GL.upload_system(A, B[None], [1.0], 1)
    W = GL._padded_system(ds, [])
    X = ds.mB[0].clone()
    info, _k = GL.chol_solve_batch([W], [X])
    assert int(info.item()) == 0
    got = X[:m, :n].cpu().numpy()
    assert np.isfinite(got).all()
    # row by row: the rows of the solution differ by many decades in the scaled cases
    num = np.abs(got - ref).max(axis=1)
    den = np.maximum(np.abs(ref).max(axis=1), 1e-300)
    assert (num / den).max() < 1e-8, (num / den).max()
    if case == "rhs_zero_rows":
        assert not got[::7].any()
    L = np.tril(W[:n, :n].cpu().numpy())
    assert np.abs(L @ L.T - A).max() <= 1e-13 * np.abs(A).max() * (1e4 if case in ("graded_rows", "diagonal_16_decades") else 1)


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 192, 448), (1024, 512, 2048)])
def test_ozaki_gemm(M, N, K):
    """b200_dev_ozaki_gemm_nt: C -= A B^T from error-free INT8 digit planes on tcgen05 (csrc/ozaki.cu) against the float64
    product (cuBLAS) on operands whose rows and entries span twelve decades: the error stays below 4e-15 of
    sum_k |a_ik| |b_jk| + |c_ij| entry by entry -- the class of a float64 dot product itself (the DMMA tile of this library
    measures 6e-15 to 2e-14 on the same inputs)."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    rnd = lambda *sh: torch.randn(sh, dtype=torch.float64, device="cuda", generator=g)  # noqa: E731
    uni = lambda *sh: torch.rand(sh, dtype=torch.float64, device="cuda", generator=g)  # noqa: E731
    A = rnd(M, K) * 10.0 ** (-6 * uni(M, 1)) * 10.0 ** (-6 * uni(M, K))
    B = rnd(N, K) * 10.0 ** (-6 * uni(N, 1))
    A[M // 2] = 0.0  # an all-zero row has scale 0
    C0 = rnd(M, N) * 1e-3
    nbytes = int(_lib.lib.b200_ozaki_gemm_work_bytes(M, N, K))
    work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    Cm = C0.clone()
    _lib.dev_ozaki_gemm_nt(GL.ptr(A), K, GL.ptr(B), K, GL.ptr(Cm), N, M, N, K, GL.ptr(work), nbytes, GL.stream_handle())
    torch.cuda.synchronize()
    ref = C0 - A @ B.T
    bound = A.abs() @ B.abs().T + C0.abs()
    assert float(((Cm - ref).abs() / bound).max()) < 4e-15
    assert torch.equal(Cm[M // 2], C0[M // 2])


def test_ozaki_gemm_extreme_row_scales():
    """Rows of A scaled by 1e+250 and rows of B by 1e-250 (and the reverse): the digit planes are taken relative to
    power-of-two row scales, so the products come out to the same relative accuracy as for rows of order 1."""
    M, N, K = 256, 128, 320
    g = torch.Generator(device="cuda").manual_seed(5)
    rnd = lambda *sh: torch.randn(sh, dtype=torch.float64, device="cuda", generator=g)  # noqa: E731
    A, B = rnd(M, K), rnd(N, K)
    A[: M // 2] *= 1e250
    A[M // 2:] *= 1e-250
    B[: N // 2] *= 1e-250
    B[N // 2:] *= 1e-20
    C0 = torch.zeros(M, N, dtype=torch.float64, device="cuda")
    nbytes = int(_lib.lib.b200_ozaki_gemm_work_bytes(M, N, K))
    work = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    Cm = C0.clone()
    _lib.dev_ozaki_gemm_nt(GL.ptr(A), K, GL.ptr(B), K, GL.ptr(Cm), N, M, N, K, GL.ptr(work), nbytes, GL.stream_handle())
    torch.cuda.synchronize()
    assert bool(torch.isfinite(Cm).all())
    bound = A.abs() @ B.abs().T
    ok = bound > 1e-300  # (products of 1e-250 rows with 1e-250 rows underflow in float64 itself)
    assert float(((Cm + A @ B.T).abs()[ok] / bound[ok]).max()) < 4e-15


def test_sliced_int8_path_matches_dmma_path():
    """The batched Cholesky + solves with the long-K updates on the INT8 tensor cores (default) and on the DMMA pipe
    (B200_OZAKI=0 / no workspace) solve the same ill-conditioned system (cond 1e6, n = 1500: three super-panels, a
    partial last one, right-hand sides with a half tile) to the same answer: both within P-f64 of LAPACK and of each
    other, and L L^T = A to 1e-13 either way."""
    from scipy.linalg import cho_solve, cholesky

    rng = np.random.default_rng(11)
    n, m = 1500, 200
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    A = (Q * np.logspace(-6, 0, n)) @ Q.T
    A = 0.5 * (A + A.T)
    B = rng.standard_normal((m, n))
    ref = cho_solve((cholesky(A, lower=True), True), B.T).T
    got = {}
    old = GL.OZAKI
    try:
        for oz in (True, False):
            GL.OZAKI = oz
            ds = GL.upload_system(A, B[None], [1.0], 1)
            W = GL._padded_system(ds, [])
            X = ds.mB[0].clone()
            info, _k = GL.chol_solve_batch([W], [X], mrows=[m])
            assert int(info.item()) == 0
            got[oz] = X[:m, :n].cpu().numpy()
            L = np.tril(W[:n, :n].cpu().numpy())
            assert rel(L @ L.T, A) < 1e-13
            assert rel(got[oz], ref) < P64
    finally:
        GL.OZAKI = old
    assert rel(got[True], got[False]) < 2 * P64  # (cond 1e6: two float64-class solvers differ by ~cond * eps)
    assert not np.array_equal(got[True], got[False])  # (two different arithmetic paths really ran)


def test_legacy_tile_path():
    """B200_TILE64=0 (the 128x128 one-CTA-per-SM tile, kept for A/B comparisons) is read once per process: run the dense
    linear-algebra tests again in a child process with it."""
    import subprocess
    import sys

    env = dict(os.environ, B200_TILE64="0")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-k",
                        "test_gemm_nt or test_chol_solve_random or test_chol_info_nonpd"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_chol_solve_batch_of_unequal_systems_with_a_failing_one():
    """One batched call over systems of different sizes (one, three and six block columns past the super-panel width,
    right-hand sides with and without a half tile) and one matrix that is NOT positive definite: every healthy system
    comes out as if solved alone, the failing one reports its LAPACK info code and stays finite."""
    from scipy.linalg import cho_solve, cholesky

    rng = np.random.default_rng(8)
    shapes = [(300, 50), (1300, 200), (900, 129), (1300, 64), (128, 128)]
    Ws, Xs, refs = [], [], []
    for k, (n, m) in enumerate(shapes):
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        A = (Q * np.logspace(-5, 0, n)) @ Q.T
        A = 0.5 * (A + A.T)
        if k == 3:
            A[700, 700] = -0.5  # not positive definite: the pivot of row 700 (or an earlier one) fails
        B = rng.standard_normal((m, n))
        ds = GL.upload_system(A, B[None], [1.0], 1)
        Ws.append(GL._padded_system(ds, []))
        Xs.append(ds.mB[0].clone())
        refs.append(None if k == 3 else cho_solve((cholesky(A, lower=True), True), B.T).T)
    info, _k = GL.chol_solve_batch(Ws, Xs)
    info = info.cpu().numpy()
    for k, (n, m) in enumerate(shapes):
        got = Xs[k][:m, :n].cpu().numpy()
        assert np.isfinite(got).all(), k
        if k == 3:
            assert 0 < info[k] <= 701
        else:
            assert info[k] == 0
            assert rel(got, refs[k]) < 1e-9, (k, rel(got, refs[k]))


def test_chol_info_nonpd():
    A = np.eye(200)
    A[150, 150] = -1.0
    ds = GL.upload_system(A, np.zeros((1, 4, 200)), [1.0], 2)
    W = GL._padded_system(ds, [])
    info, _k = GL.chol_solve_batch([W], None)
    assert int(info.item()) == 151  # LAPACK dpotrf convention


def test_eigh_device():
    rng = np.random.default_rng(3)
    n = 300
    Gm = rng.standard_normal((n, n))
    Q, _ = np.linalg.qr(Gm)
    lam_true = np.concatenate([np.zeros(60), np.logspace(-8, 0, n - 60)])
    A = (Q * lam_true) @ Q.T
    A = 0.5 * (A + A.T)
    ds = GL.upload_system(A, np.zeros((1, 4, n)), [1.0], 2)
    lam, Vt, sweeps = GL.eigh_device(ds.A.clone(), n)
    lam, V = lam[:n].cpu().numpy(), Vt[:n, :n].cpu().numpy().T
    # absolute accuracy class of LAPACK's eigh: a small multiple of sqrt(n) eps |A|_F (|A|_F ~ 5 here)
    assert np.abs(np.sort(lam) - np.linalg.eigvalsh(A)).max() < 5e-13
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12
    assert np.abs(A @ V - V * lam).max() < 5e-13
    assert 0 <= sweeps < 40  # (0: the tridiagonalisation-based solver; the Jacobi solver counts its sweeps)


def _graded_matrix(n, seed):
    """Spectrum like the system matrices of this path: a numerically zero cluster + ten decades, ~uniform in log."""
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.zeros(n // 5), np.logspace(-11, -1, n - n // 5)])
    A = (Q * lam) @ Q.T
    return 0.5 * (A + A.T)


def _padded(A):
    n = A.shape[0]
    npad = (n + 127) // 128 * 128
    Ad = torch.eye(npad, dtype=torch.float64, device="cuda")
    Ad[:n, :n] = torch.from_numpy(A).cuda()
    return Ad


def test_eigh_batch_of_unequal_sizes():
    """The batched eigensolver (csrc/trieig.cu) on systems of different sizes in ONE call, including sizes around the
    panel width of the blocked tridiagonalisation (64) and the compact-WY panel (128) and the trivial ones."""
    sizes = [1, 2, 3, 63, 64, 65, 66, 127, 129, 191, 200, 333]
    mats = [_graded_matrix(n, 10 + n) if n > 3 else np.random.default_rng(n).standard_normal((n, n)) for n in sizes]
    mats = [0.5 * (A + A.T) for A in mats]
    res, _ = GL.eigh_device_batch([(_padded(A), A.shape[0]) for A in mats])
    for A, (lam, Vt) in zip(mats, res):
        n = A.shape[0]
        npad = Vt.shape[0]
        lam, V = lam[:n].cpu().numpy(), Vt[:n, :n].cpu().numpy().T
        scale = max(np.abs(A).max(), 1e-300)
        assert np.abs(np.sort(lam) - np.linalg.eigvalsh(A)).max() < 1e-13 * scale * max(n, 10), n
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-13, n
        assert np.abs(A @ V - V * lam).max() < 1e-13 * scale * max(n, 10), n
        # the padding stays the identity and does not leak into the vectors
        assert torch.equal(Vt[n:, n:], torch.eye(npad - n, dtype=torch.float64, device="cuda"))
        assert float(Vt[:n, n:].abs().max()) == 0.0 if npad > n else True


def test_tridiag_blocked_equals_unblocked_spectrum():
    """b200_dev_tridiag: the tridiagonal matrix has the spectrum of A (blocked dlatrd-style reduction, panels of 64)."""
    import scipy.linalg as sl

    for n in (5, 64, 130, 400):
        A = _graded_matrix(n, n) if n > 5 else np.diag(np.arange(1.0, 6.0)) + 0.1
        Ad = _padded(A)
        d, e, tau = (torch.zeros(n, dtype=torch.float64, device="cuda") for _ in range(3))
        _lib.dev_tridiag(GL.ptr(Ad), Ad.stride(0), n, GL.ptr(d), GL.ptr(e), GL.ptr(tau), GL.stream_handle())
        torch.cuda.synchronize()
        lt = sl.eigvalsh_tridiagonal(d.cpu().numpy(), e.cpu().numpy()[: n - 1])
        assert np.abs(lt - np.linalg.eigvalsh(A)).max() < 1e-14 * max(np.abs(A).max(), 1.0) * n


def _nasty_matrices():
    rng = np.random.default_rng(77)

    def rot(lam):
        n = len(lam)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        A = (Q * np.asarray(lam, dtype=np.float64)) @ Q.T
        return 0.5 * (A + A.T)

    n = 150
    w = np.abs(np.arange(21) - 10.0)  # Wilkinson W21+: pairs of eigenvalues agreeing to ~1e-14
    W21 = np.diag(w) + np.diag(np.ones(20), 1) + np.diag(np.ones(20), -1)
    return {
        "zero": np.zeros((40, 40)),
        "identity": np.eye(70),
        "diagonal_distinct": np.diag(np.linspace(-3.0, 5.0, 90)),
        "diagonal_repeats": np.diag(np.repeat([1.0, 2.0, 2.0, 7.0, 7.0, 7.0], 20)),
        "rank_one": np.outer(np.arange(1.0, 81.0), np.arange(1.0, 81.0)),
        "exact_multiplicities": rot(np.repeat([0.0, 1e-9, 1.0, 1.0 + 1e-13, 3.0], n // 5)),
        "wilkinson": W21,
        "wilkinson_embedded": rot(np.concatenate([np.linalg.eigvalsh(W21), np.linspace(20, 30, 60)])),
        "tiny_scale": 1e-200 * rot(np.logspace(-8, 0, n)),
        "huge_scale": 1e150 * rot(np.logspace(-8, 0, n)),
        "negative_definite": -rot(np.logspace(-12, 0, n)),
        "indefinite_graded": rot(np.concatenate([-np.logspace(-14, -2, n // 2), np.logspace(-14, 0, n - n // 2)])),
    }


def test_eigh_on_degenerate_and_badly_scaled_matrices():
    """The eigensolver must return a finite orthonormal eigenbasis for every symmetric input: exact multiplicities,
    reducible (diagonal) matrices, the zero matrix, Wilkinson pairs, scales near the ends of the float64 range.
    (Whether it gets there through inverse iteration or through the Jacobi fallback is reported, not asserted.)"""
    mats = _nasty_matrices()
    before = _lib.eigh_fallback_count()
    res, _ = GL.eigh_device_batch([(_padded(A), A.shape[0]) for A in mats.values()])
    print("fallbacks:", _lib.eigh_fallback_count() - before, "of", len(mats))
    for (name, A), (lam, Vt) in zip(mats.items(), res):
        n = A.shape[0]
        lam, V = lam[:n].cpu().numpy(), Vt[:n, :n].cpu().numpy().T
        assert np.isfinite(lam).all() and np.isfinite(V).all(), name
        scale = np.abs(A).max()
        tol = 1e-13 * scale * n + 1e-299  # (bisection brackets are never narrower than its 1e-300 pivot floor)
        assert np.abs(np.sort(lam) - np.linalg.eigvalsh(A)).max() <= tol, name
        assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12, name
        assert np.abs(A @ V - V * lam).max() <= tol, name


def test_eigh_falls_back_to_jacobi_when_orthonormalisation_fails(monkeypatch):
    """B200_EIGH_TEST_FAIL gives every inverse-iteration thread the same shift: the vectors are linearly dependent, the
    Gram matrix of the Cholesky-QR stage is singular, and the system must be solved again by the Jacobi solver - never
    returned as garbage."""
    n = 200
    A = _graded_matrix(n, 5)
    before = _lib.eigh_fallback_count()
    monkeypatch.setenv("B200_EIGH_TEST_FAIL", "1")
    lam, Vt, _ = GL.eigh_device(_padded(A), n)
    monkeypatch.delenv("B200_EIGH_TEST_FAIL")
    assert _lib.eigh_fallback_count() == before + 1
    lam, V = lam[:n].cpu().numpy(), Vt[:n, :n].cpu().numpy().T
    assert np.isfinite(V).all()
    assert np.abs(np.sort(lam) - np.linalg.eigvalsh(A)).max() < 1e-13
    assert np.abs(V.T @ V - np.eye(n)).max() < 1e-12
    assert np.abs(A @ V - V * lam).max() < 1e-13
    # and without the hook the fast path does not fall back
    GL.eigh_device(_padded(A), n)
    assert _lib.eigh_fallback_count() == before + 1


# ---------------------------------------------------------------------------------------------------
# kernel-class seam on the reference's own unit-test inputs (tests/pyimcom/test_la.py)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.LA_CASES))
def test_la_kernels(name, golden_dir):
    g = np.load(os.path.join(golden_dir, "la.npz"))
    kern, kappaC, extra = cases.LA_CASES[name]
    outst = cases.la_outst(kappaC, **extra)
    K = getattr(GL, kern)(outst)
    K(keep_f64=True)
    ref = cases.la_outst(kappaC, **extra)
    KO = getattr(OL, kern)(ref)
    KO()
    tolT = 1e-4 if kern == "IterKernel" else 2e-5  # tiny 6x6 singular system: cond ~ 1/kappa amplifies rounding
    assert rel(outst.T, g[name + "_T"]) < tolT
    assert np.abs(outst.UC - g[name + "_UC"]).max() < tolT
    assert rel(outst.Sigma, g[name + "_Sigma"]) < tolT
    assert rel(outst.kappa, g[name + "_kappa"]) < 2e-6
    assert rel(outst.T, ref.T) < tolT
    UC, Sig, kap = outst.UC.ravel(), outst.Sigma.ravel(), outst.kappa.ravel()
    if name == "eigen3":  # test_la.py:146-159
        for j in range(16):
            assert (UC[j] < 1e-4 and 5e-4 < kap[j] < 1.5e-3) if j % 5 == 0 else (0.05 < UC[j] < 0.2 and 5e-6 < kap[j] < 1.5e-5)
            assert 0.6 < Sig[j] < 1.0
    if name == "iter2":  # test_la.py:221-230
        for j in range(16):
            assert (UC[j] < 1e-4 and 2e-3 < kap[j] < 4e-3) if j % 5 == 0 else (0.05 < UC[j] < 0.2 and 2e-4 < kap[j] < 4e-4)
    if kern == "IterKernel":
        assert np.array_equal(K.f64[0]["niter"].ravel(), KO.f64[0]["niter"].ravel())  # P-discrete


def test_incr_repair():
    """tests/pyimcom/test_la.py:8-24: the eigen-shift repair of a non-PD matrix, through the GPU CholKernel."""
    N = 6
    idx = np.arange(N)
    d = 2 * np.pi * (idx[:, None] - idx[None, :]) / N
    A = sum(np.cos(k * d) / k / N for k in range(1, N // 2 + 1)) - 1e-3 * np.identity(N)
    ds = GL.upload_system(A, np.eye(N)[None], [1.0], 2)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        (X,) = GL._chol_with_repair(ds, [[1e-4]], 0)
    assert any("repaired" in str(w.message) for w in wlist)
    Minv = X[:N, :N].cpu().numpy()  # (A + 1e-4 I + shift I)^-1
    w = np.linalg.eigvalsh(np.linalg.inv(Minv))
    assert abs(w[0] - 1e-4) < 1e-7


# ---------------------------------------------------------------------------------------------------
# whole OutStamp path on the seeded synthetic blocks
# ---------------------------------------------------------------------------------------------------
KERN = {"Cholesky": OL.CholKernel, "Eigen": OL.EigenKernel, "Iterative": OL.IterKernel, "Empirical": OL.EmpirKernel}


@pytest.mark.parametrize("name", list(cases.BLOCK_CASES))
def test_block_vs_oracle_and_reference(name, golden_dir):
    spec = cases.BLOCK_CASES[name]
    g = np.load(os.path.join(golden_dir, f"block_{name}.npz"))
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)  # PSF sampling through the GPU function seam
    gb = GpuBlock(blk, tab).prepare(stamps=spec["stamps"])
    otab = PSFTables(blk, R.iD5512C, R.gridD5512C)
    for (j, i) in spec["stamps"]:
        tag = f"s{j}_{i}_"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            s = GpuOutStamp(gb, j, i)
        o = OracleOutStamp(blk, otab, j, i)
        o.build_system_matrices()
        assert np.array_equal(s.inpix_cumsum, g[tag + "inpix_cumsum"])
        # stage (a): P-f64 against the oracle and, where stored, against the reference itself
        assert rel(s.sysmata, o.sysmata) < P64
        assert rel(s.mhalfb, o.mhalfb) < P64
        assert np.array_equal(s.sysmata, s.sysmata.T)
        if spec.get("store_ab"):
            assert rel(s.sysmata, g[tag + "sysmata"]) < P64
            assert rel(s.mhalfb, g[tag + "mhalfb"]) < P64
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            k = KERN[spec["kernel"]](o)
            k()
        kern = spec["kernel"]
        if kern == "Cholesky" and name != "repair":
            for jo in range(blk.cfg.n_out):
                assert rel(s.Ti64[jo], k.f64[jo]["Ti"]) < P64  # P-f64 on the pre-cast solution
            if spec.get("store_ti64"):
                assert rel(s.Ti64[0], g[tag + "Ti64"][0]) < P64
        if kern == "Cholesky" and len(spec["kappaC"]) > 1:
            assert np.array_equal(s.extras[0]["iv"], k.f64[0]["iv"])  # P-discrete
            assert np.array_equal(s.extras[0]["branch"], k.f64[0]["branch"])
        if kern == "Iterative":
            assert np.array_equal(s.extras[0]["niter"].ravel(), k.f64[0]["niter"].ravel())  # P-discrete: CG iterations
        o.post_kernel()
        o.perform_coaddition()
        tol = P32
        if kern == "Eigen":
            tol = 2e-5  # 1/(lam+kappa) amplification at kappa/C = 1e-5 (see tests/test_oracle_golden.py)
        if kern == "Iterative":
            tol = 1e-3  # CG stopped at rtol = 1.5e-3 (see tests/test_oracle_golden.py)
        if name == "repair":
            tol = 2e-5  # the shift |w0| is an eigenvalue of A: LAPACK vs Jacobi differ by O(eps |A|), amplified by 1/kappa
        for nm in ("T", "Sigma", "kappa", "outimage", "Tsum_stamp", "Tsum_inpix", "Neff"):
            t = tol if nm == "T" or tol < 1e-3 else 5 * tol
            assert rel(getattr(s, nm), g[tag + nm]) < t, (nm, "vs reference")
            assert rel(getattr(s, nm), getattr(o, nm)) < t, (nm, "vs oracle")
        assert np.abs(s.UC - g[tag + "UC"]).max() < tol * max(1.0, np.abs(g[tag + "UC"]).max())


@pytest.mark.parametrize("name", ["pad4", "nout2split", "chol1"])
def test_pair_block_cache_is_bit_identical(name):
    """A cut from cached InStamp-pair blocks (device SysMatA) == A from the fused per-stamp kernel, bit for bit; the
    cache interpolates fewer entries than the stamps hold, and survives a pool too small for more than one stamp."""
    spec = cases.BLOCK_CASES[name]
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    stamps = list(blk.stamp_order())
    fused = GpuBlock(blk, tab, a_cache=False).prepare(stamps=stamps)
    cached = GpuBlock(blk, tab, a_cache=True).prepare(stamps=stamps)
    tiny = GpuBlock(blk, tab, a_cache=True)
    tiny.pool_bytes = 8  # forces an eviction + regrow on every request
    tiny.prepare(stamps=stamps)
    tot = 0
    for k in range(len(stamps)):
        if fused.plans[stamps[k]].n == 0:
            continue
        a0 = fused.build_system(k)[0].A
        a1 = cached.build_system(k)[0].A
        a2 = tiny.build_system(k)[0].A
        assert torch.equal(a0, a1) and torch.equal(a0, a2)
        tot += fused.plans[stamps[k]].n ** 2 / 2
    if len(stamps) > 1:
        assert cached.pair_points < tot


@pytest.mark.parametrize("name", ["pad4", "nout2split"])
def test_pair_blocks_resident_grid_is_bit_identical(name, monkeypatch):
    """The experiment form of the pair-block kernel (B200_PAIR_RESIDENT: a few resident CTAs per SM walking the tile
    list, csrc/interp.cu) writes the same pool, bit for bit, as one CTA per tile -- self blocks with their mirrored
    lower tiles included."""
    spec = cases.BLOCK_CASES[name]
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    stamps = list(blk.stamp_order())
    gb = GpuBlock(blk, tab, a_cache=True).prepare(stamps=stamps)
    plans = [gb.plans[s] for s in stamps if gb.plans[s].n > 0]
    monkeypatch.delenv("B200_PAIR_RESIDENT", raising=False)
    gb.reset_cache()
    gb.ensure_pairs(plans)  # (allocates the pool)
    used = gb._pool_used
    assert used > 0
    gb._pool[:used].fill_(float("nan"))
    gb.reset_cache()
    gb.ensure_pairs(plans)
    torch.cuda.synchronize()
    assert gb._pool_used == used
    ref = gb._pool[:used].clone()
    assert not bool(torch.isnan(ref).all())
    for per_sm in ("1", "3"):
        monkeypatch.setenv("B200_PAIR_RESIDENT", per_sm)
        gb._pool[:used].fill_(float("nan"))
        gb.reset_cache()
        gb.ensure_pairs(plans)
        torch.cuda.synchronize()
        assert gb._pool_used == used
        # (padding columns of a block row are never written by either form: compare where the reference form wrote)
        got = gb._pool[:used]
        assert torch.equal(torch.nan_to_num(got, nan=0.0), torch.nan_to_num(ref, nan=0.0))


@pytest.mark.parametrize("name", ["chol1", "pad4", "nout2split", "amp", "oversamp7", "oversamp5_3img", "one_image"])
def test_device_tables(name):
    """PSF-overlap tables built on the device (device iD5512C sampling + partial DFTs as DMMA
    products) equal the NumPy-FFT tables of the host builder, and a block coadded from them equals the oracle."""
    from pyimcom_b200.psfovl_device import DeviceTables

    if name == "amp":  # amplitude penalty on the Fourier modes (psfutil.py:661-671) + circular cut + normalisation
        spec = dict(cases.BLOCK_CASES["chol1"], cfg=dict(cases._MINI, amp_penalty=(0.5, 0.8), psf_circ=True, psf_norm=True))
    elif name == "oversamp7":  # sizes with no factor of two anywhere: ns, nfft and the kept lags are odd
        spec = dict(cases.BLOCK_CASES["chol1"], cfg=dict(cases._MINI, oversamp=7, npixpsf=14))
    elif name == "oversamp5_3img":
        spec = dict(cases.BLOCK_CASES["chol1"], cfg=dict(cases._MINI, oversamp=5, npixpsf=20), n_image=3)
    elif name == "one_image":
        spec = dict(cases.BLOCK_CASES["chol1"], n_image=1)
    else:
        spec = cases.BLOCK_CASES[name]  # "nout2split": PSF splitting, kept lags 2 ns + 1, two output PSFs
    blk = cases.make_block(spec)
    host = PSFTables(blk, G.iD5512C, G.gridD5512C)
    dev = DeviceTables(blk, G.iD5512C, G.gridD5512C)
    assert rel(dev.outovlc, host.outovlc) < 1e-12
    groups = sorted({(j >> 1 << 1, i >> 1 << 1) for j in range(blk.cfg.n1P + 2) for i in range(blk.cfg.n1P + 2)})
    for Gk in groups:
        host.group(Gk)
        dev.group(Gk)
        assert host.grp_imgs[Gk] == dev.grp_imgs[Gk]
        if not host.grp_imgs[Gk]:
            continue
        assert rel(dev.get_self(Gk).cpu().numpy(), host.get_self(Gk)) < 1e-11
        assert rel(dev.get_io(Gk).cpu().numpy(), host.get_io(Gk)) < 1e-11
    for Ga in groups:
        for Gb in groups:
            if Ga < Gb and host.grp_imgs[Ga] and host.grp_imgs[Gb] and max(abs(Ga[0] - Gb[0]), abs(Ga[1] - Gb[1])) <= 2:
                assert rel(dev.get_cross(Ga, Gb).cpu().numpy(), host.get_cross(Ga, Gb)) < 1e-11
    gb = GpuBlock(blk, dev).prepare(stamps=spec["stamps"])
    otab = PSFTables(blk, R.iD5512C, R.gridD5512C)
    j, i = spec["stamps"][0]
    s = GpuOutStamp(gb, j, i)
    o = OracleOutStamp(blk, otab, j, i)
    o.build_system_matrices()
    OL.CholKernel(o)()
    o.post_kernel()
    o.perform_coaddition()
    assert rel(s.sysmata, o.sysmata) < P64 and rel(s.mhalfb, o.mhalfb) < P64
    assert rel(s.T, o.T) < P32 and rel(s.outimage, o.outimage) < 5 * P32


def test_block_run_maps():
    """Whole-block loop: the accumulated maps equal the overlap-add of the per-stamp oracle results."""
    spec = cases.BLOCK_CASES["pad4"]
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    gb = GpuBlock(blk, tab).prepare()
    n0 = _lib.launch_count()
    gb.run()
    maps = gb.download()
    assert _lib.launch_count() > n0
    cfg = blk.cfg
    otab = PSFTables(blk, R.iD5512C, R.gridD5512C)
    side = cfg.NsideP + 2 * cfg.fade_kernel
    out = np.zeros((cfg.n_out, cfg.n_inframe, side, side), dtype=np.float32)
    UC = np.zeros((cfg.n_out, side, side), dtype=np.float32)
    for (j, i) in blk.stamp_order():
        o = OracleOutStamp(blk, otab, j, i)
        o.build_system_matrices()
        OL.CholKernel(o)()
        o.post_kernel()
        o.perform_coaddition()
        b, l = (j - 1) * cfg.n2, (i - 1) * cfg.n2
        out[:, :, b:b + cfg.n2f, l:l + cfg.n2f] += o.outimage
        UC[:, b:b + cfg.n2f, l:l + cfg.n2f] += o.UC
    assert rel(maps["out_map"], out) < 5e-6
    assert rel(maps["UC_map"], UC) < 5e-6


@pytest.mark.parametrize("nfr", [9, 19])
def test_many_input_layers(nfr):
    """More input layers than one DMMA pass (8) and than one T-apply launch (16) takes: every coadded layer equals the
    oracle's (the reference's einsum over any number of layers, coadd.py:1339-1350)."""
    spec = dict(cases.BLOCK_CASES["chol1"])
    spec["cfg"] = dict(spec["cfg"], n_inframe=nfr)
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    j, i = spec["stamps"][0]
    s = GpuOutStamp(GpuBlock(blk, tab).prepare(stamps=[(j, i)]), j, i)
    o = OracleOutStamp(blk, PSFTables(blk, R.iD5512C, R.gridD5512C), j, i)
    o.build_system_matrices()
    OL.CholKernel(o)()
    o.post_kernel()
    o.perform_coaddition()
    assert s.outimage.shape == o.outimage.shape and s.outimage.shape[1] == nfr
    for f in range(nfr):
        assert rel(s.outimage[:, f], o.outimage[:, f]) < 5e-6, f
    assert rel(s.T, o.T) < 2e-6


def test_strip_sharded_block_equals_unsharded():
    """SURVEY 8e, single-block sharding: the block coadded as two strips of 2x2 stamp-group rows (shard.
    assign_stamp_groups; each strip adds into its own zero-initialised cube, the cubes are summed as shard.reduce_cube
    does) equals the unsharded block bit for bit except on the rows where the fade borders of the two strips overlap,
    and there to float32 rounding of the seam adds.  (bench.py --strong runs the same check across real ranks.)"""
    from pyimcom_b200.shard import assign_stamp_groups

    spec = cases.BLOCK_CASES["pad4"]
    blk = cases.make_block(spec)
    cfg = blk.cfg
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    ref = GpuBlock(blk, tab).prepare().run().download()
    world = 2
    parts = [GpuBlock(blk, tab).prepare(stamps=assign_stamp_groups(cfg.n1P, world, r)).run().download() for r in range(world)]
    assert sorted(sum((assign_stamp_groups(cfg.n1P, world, r) for r in range(world)), [])) == sorted(blk.stamp_order())
    y = 2 * ((len(range(1, cfg.n1P + 1, 2)) + world - 1) // world) * cfg.n2
    seam = np.zeros(ref["out_map"].shape[-2], dtype=bool)
    seam[y:y + 2 * cfg.fade_kernel] = True
    for k in ("out_map", "UC_map", "Sigma_map", "kappa_map", "Tsum_map", "Neff_map"):
        tot = parts[0][k] + parts[1][k]
        assert np.array_equal(tot[..., ~seam, :], ref[k][..., ~seam, :]), k
        assert rel(tot, ref[k]) < 5e-7, k
    assert np.array_equal(parts[0]["T_weightmap"] + parts[1]["T_weightmap"], ref["T_weightmap"])


def test_empirical_without_quality_control():
    """cfg.no_qlt_ctrl with the empirical kernel (coadd.py:856-858, 1020-1025; lakernel.py:770-774): no system matrix is
    interpolated (no pair-block / mBhalf launch), T and therefore the coadded image are those of the regular empirical
    run, and the U/C, Sigma and kappa maps stay at zero as in the reference."""
    spec = cases.BLOCK_CASES["empir"]
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    ref = GpuBlock(blk, tab).prepare().run().download()
    blk.cfg.no_qlt_ctrl = True
    try:
        gb = GpuBlock(blk, tab).prepare()
        _lib.profile(1)
        gb.run()
        torch.cuda.synchronize()
        kinds = set(_lib.profile_read())
        _lib.profile(0)
        got = gb.download()
    finally:
        blk.cfg.no_qlt_ctrl = False
    assert not kinds & {"build_A", "build_B", "assemble_A"}, kinds
    assert np.array_equal(got["out_map"], ref["out_map"]) and np.array_equal(got["Tsum_map"], ref["Tsum_map"])
    assert np.array_equal(got["Neff_map"], ref["Neff_map"])
    for k in ("UC_map", "Sigma_map", "kappa_map"):
        assert not got[k].any() and ref[k].any(), k


def test_pipelined_batches_equal_single_batch():
    """GpuBlock.run(): four pipelined batches of 4 stamps (stage (a) of batch k+1 and the T-apply of batch k overlap the
    factorisation) give bit-identical block maps to one batch of 16, with and without the concurrent solve streams."""
    spec = cases.BLOCK_CASES["pad4"]
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    ref = GpuBlock(blk, tab).prepare()
    ref.run(batch=16)
    want = ref.download()
    for streams, tiny_pool in ((3, False), (1, False), (3, True)):
        old = GL.SOLVE_STREAMS
        GL.SOLVE_STREAMS = streams
        try:
            gb = GpuBlock(blk, tab)
            if tiny_pool:
                gb.pool_bytes = 8  # every batch evicts the pair-block cache while the previous batch is still in flight
            gb.prepare()
            gb.run(batch=4)
            got = gb.download()
        finally:
            GL.SOLVE_STREAMS = old
        for k in want:
            assert np.array_equal(got[k], want[k]), (k, streams, tiny_pool)


@pytest.mark.parametrize("kernel,kappaC", [("Eigen", [1e-5, 1e-4, 1e-3]), ("Eigen", [5e-4]), ("Iterative", [1e-2])])
def test_eigen_and_iter_blocks_batched_equal_stamp_by_stamp(kernel, kappaC):
    """GpuBlock.run() with the Eigen / Iterative kernels: the whole block in batches of 16 (one batched eigendecomposition
    over stamps of different sizes) against the same block coadded one stamp per batch."""
    spec = dict(cases.BLOCK_CASES["pad4"], kernel=kernel, kappaC=kappaC)
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    one = GpuBlock(blk, tab).prepare()
    one.run(batch=1)
    want = one.download()
    gb = GpuBlock(blk, tab).prepare()
    gb.run(batch=16)
    got = gb.download()
    for k in want:
        scale = max(np.abs(want[k]).max(), 1e-30)
        assert np.abs(got[k].astype(np.float64) - want[k]).max() <= 5e-6 * scale, k


@pytest.mark.parametrize("kernel,kappaC,tol", [("Eigen", [1e-5, 1e-4, 1e-3], 2e-5), ("Eigen", [5e-4], 2e-5),
                                               ("Iterative", [1e-2], 1e-3), ("Cholesky", [1e-5, 1e-4, 1e-3], 2e-6)])
def test_two_output_psfs_with_every_kernel(kernel, kappaC, tol):
    """n_out = 2 with PSF splitting and a flat penalty (the nout2split block) through the Eigen, Iterative and multi-kappa
    Cholesky kernels: one decomposition / one system matrix serves both output PSFs (lakernel.py:154-223, 281-394,
    592-744); every output PSF against the oracle."""
    spec = dict(cases.BLOCK_CASES["nout2split"], kernel=kernel, kappaC=kappaC)
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    j, i = spec["stamps"][0]
    s = GpuOutStamp(GpuBlock(blk, tab).prepare(stamps=[(j, i)]), j, i)
    o = OracleOutStamp(blk, PSFTables(blk, R.iD5512C, R.gridD5512C), j, i)
    o.build_system_matrices()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        {"Eigen": OL.EigenKernel, "Iterative": OL.IterKernel, "Cholesky": OL.CholKernel}[kernel](o)()
    o.post_kernel()
    o.perform_coaddition()
    assert s.T.shape[0] == 2
    for j_out in range(2):
        assert rel(s.T[j_out], o.T[j_out]) < tol, j_out
        assert rel(s.outimage[j_out], o.outimage[j_out]) < 5 * tol, j_out
        assert rel(s.Sigma[j_out], o.Sigma[j_out]) < 5 * tol and rel(s.kappa[j_out], o.kappa[j_out]) < 5 * tol, j_out


@pytest.mark.parametrize("over,n_image", [(dict(oversamp=7, npixpsf=14), 2), (dict(oversamp=5, npixpsf=20), 3),
                                          (dict(), 1), (dict(), 7), (dict(fade_kernel=0), 2), (dict(fade_kernel=3), 2),
                                          (dict(n2=9), 2)])
def test_unusual_geometries(over, n_image):
    """Shapes the other cases do not visit: oversampling factors without a polyphase specialisation (7) and with one (5),
    a single input image, seven input images, no fade margin, a fade margin of 3, an odd stamp size.  A, -B/2, T and the
    coadded layers of one interior stamp against the oracle, through the cached pair blocks (GpuBlock) and the fused
    assembly (a_cache off)."""
    spec = dict(cases.BLOCK_CASES["pad4"])
    spec["cfg"] = dict(spec["cfg"], **over)
    spec["n_image"] = n_image
    blk = cases.make_block(spec)
    j, i = 2, 2
    o = OracleOutStamp(blk, PSFTables(blk, R.iD5512C, R.gridD5512C), j, i)
    o.build_system_matrices()
    OL.CholKernel(o)()
    o.post_kernel()
    o.perform_coaddition()
    for a_cache in (True, False):
        gb = GpuBlock(blk, PSFTables(blk, G.iD5512C, G.gridD5512C))
        gb.a_cache = a_cache
        s = GpuOutStamp(gb.prepare(stamps=[(j, i)]), j, i)
        assert np.array_equal(np.asarray(s.inpix_cumsum), np.asarray(o.inpix_cumsum))
        assert rel(s.sysmata, o.sysmata) < 1e-9 and rel(s.mhalfb, o.mhalfb) < 1e-9, a_cache
        assert rel(s.T, o.T) < 2e-6 and rel(s.outimage, o.outimage) < 1e-5, a_cache


def test_repair_branch_survives_pool_eviction():
    """The eigen-shift repair of CholKernel._cholesky_wrapper (lakernel.py:262-279) re-assembles A from the cached
    InStamp-pair blocks.  In the pipelined run the blocks of batch k+1 are requested before batch k is finished; with a
    pool that has to be evicted for every batch the pending batch must be completed first (GpuBlock.ensure_pairs'
    before_evict), otherwise the repair would read overwritten blocks: maps must equal the single-batch run bit for bit."""
    spec = cases.BLOCK_CASES["repair"]
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    with warnings.catch_warnings(record=True) as w0:
        warnings.simplefilter("always")
        ref = GpuBlock(blk, tab).prepare()
        ref.run(batch=16)
        want = ref.download()
    n_rep = sum("repaired" in str(w.message) for w in w0)
    assert n_rep > 0  # the case is built to fire the branch
    with warnings.catch_warnings(record=True) as w1:
        warnings.simplefilter("always")
        from pyimcom_b200.coadd import release_parked_pools

        release_parked_pools()  # (a pool parked by an earlier, larger block would be big enough never to evict)
        gb = GpuBlock(blk, tab)
        gb.pool_bytes = 8  # every batch evicts
        gb.prepare()
        gb.run(batch=4)
        got = gb.download()
    assert gb.pool_evictions >= 2
    assert sum("repaired" in str(w.message) for w in w1) == n_rep
    for k in want:
        assert np.array_equal(got[k], want[k]), k


@pytest.mark.parametrize("kern", ["CholKernel", "EigenKernel", "IterKernel", "EmpirKernel"])
def test_kernel_without_input_pixels(kern):
    """lakernel.py:110-119: a postage stamp with no input pixels gives an empty T, U/C = kappa = 1, Sigma = 0."""
    outst = cases.la_outst([1e-2], iterative=True)
    outst.inpix_cumsum = np.array([0])
    outst.sysmata = np.zeros((0, 0))
    outst.mhalfb = np.zeros((1, 16, 0))
    outst.iny_val = outst.inx_val = np.zeros(0)
    getattr(GL, kern)(outst)()
    assert outst.T.shape == (1, 16, 0) and outst.T.dtype == np.float32
    assert np.all(outst.UC == 1) and np.all(outst.kappa == 1) and np.all(outst.Sigma == 0)
    assert outst.UC.shape == (1, 4, 4) and outst.UC.dtype == np.float32


# ---------------------------------------------------------------------------------------------------
# block output assembly
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.OUTPUT_CASES))
def test_output_assembly(name, golden_dir):
    """Device un-fade + crop + log-integer encoding against the oracle and against the arrays the reference itself
    produced (tests/golden/output.npz).  P-f32 here is exact (same float64 division rounded to float32); the 16-bit
    codes are identical up to one count on <= 0.5 % of the pixels (float32 log10 of the host, cases.codes_match)."""
    from oracle import output as OO
    from pyimcom_b200.coadd import assemble_output

    g = np.load(os.path.join(golden_dir, "output.npz"))
    cfg, maps, n_inimage, pad_sides, is_final = cases.output_case(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        enc_in = {}
        want = OO.build_output(maps, cfg, n_inimage, is_final, pad_sides, keep_inputs=enc_in)
    dmaps = {k: torch.from_numpy(v).cuda() for k, v in maps.items()}
    got = assemble_output(dmaps, cfg, n_inimage, is_final, pad_sides)
    assert all(torch.equal(dmaps[k].cpu(), torch.from_numpy(maps[k])) for k in maps)  # block maps untouched
    assert set(got) == set(want)
    for k in ("PRIMARY", "INWEIGHT", "INWTFLAT"):
        assert got[k].dtype == np.float32 and np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["PRIMARY"], g[name + "_PRIMARY"])
    for e in set(got) - {"PRIMARY", "INWEIGHT", "INWTFLAT"}:
        assert cases.codes_match(got[e], want[e]), e
        assert cases.codes_match(got[e], g[name + "_" + e]), e
        # integer output: every code that differs from the reference's is a tie of the reference's own float32 formula
        x, coef = enc_in[e]
        nm_o, nt = cases.code_mismatches_are_ties(got[e], want[e], x, coef)
        nm_r, _ = cases.code_mismatches_are_ties(got[e], g[name + "_" + e], x, coef)
        print(f"{name} {e}: {nm_r} of {x.size} codes differ from the reference ({nm_o} from this host's NumPy), "
              f"all inside the float32 ambiguity band ({nt} pixels)")


def test_output_assembly_of_a_block(golden_dir):
    """GpuBlock.build_output on the maps of a coadded block equals the oracle's assembly of the downloaded maps."""
    from oracle import output as OO

    spec = cases.BLOCK_CASES["chol1"]
    blk = cases.make_block(spec)
    tab = PSFTables(blk, G.iD5512C, G.gridD5512C)
    gb = GpuBlock(blk, tab).prepare().run()
    maps = gb.download()
    got = gb.build_output(is_final=True, pad_sides="")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        enc_in = {}
        want = OO.build_output(maps, blk.cfg, blk.n_inimage, True, "", keep_inputs=enc_in)
    assert np.array_equal(got["PRIMARY"], want["PRIMARY"]) and np.array_equal(got["INWTFLAT"], want["INWTFLAT"])
    for e in ("FIDELITY", "SIGMA", "KAPPA", "INWTSUM", "EFFCOVER"):
        assert cases.codes_match(got[e], want[e], max_frac=0.01), e
        cases.code_mismatches_are_ties(got[e], want[e], *enc_in[e])
    assert np.array_equal(gb.download()["out_map"], maps["out_map"])


# ---------------------------------------------------------------------------------------------------
# input-pixel partitioning
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.PARTITION_CASES))
def test_partition(name, golden_dir):
    """Device binning + layer extraction: identical (index work) to the oracle and to the reference's own arrays."""
    from oracle import partition as OP
    from pyimcom_b200.partition import DevicePartition, to_host

    g = np.load(os.path.join(golden_dir, "partition.npz"))
    cfg, outpix, mask_a, mask_b, use, indata, sca, sp_res = cases.partition_case(name)
    dp = DevicePartition(cfg, use, sca_nside=sca, sp_res=sp_res)
    got = to_host(dp.partition(outpix, mask_a & mask_b, indata))
    assert got["is_relevant"] == bool(g[name + "_is_relevant"])
    if not got["is_relevant"]:
        return
    want = OP.partition_pixels(outpix, mask_a & mask_b, cfg, use, sca_nside=sca, sp_res=sp_res)
    assert dp.npixmax == want["y_idx"].shape[-1]
    for k in ("pix_count", "y_idx", "x_idx", "y_val", "x_val"):
        assert got[k].dtype == want[k].dtype, (k, got[k].dtype)
        assert np.array_equal(got[k], want[k]), k
        assert np.array_equal(got[k], g[name + "_" + k]), k
    assert got["max_count"] == want["max_count"]
    assert np.array_equal(got["data"], g[name + "_data"])


def test_partition_large_cells_and_overflow():
    """Default-sized cells (45 x 45 detector pixels: several 256-pixel chunks per cell, many pixels per stamp and cell) on
    a 4088-pixel detector, against the oracle; and the reference's overflow (IndexError) when npixmax is too small."""
    from oracle import partition as OP
    from pyimcom_b200.partition import DevicePartition, to_host
    from pyimcom_b200.synth import StampConfig

    cfg = StampConfig(n1=2, n2=40, postage_pad=1, dtheta_arcsec=0.04, fade_kernel=1)
    rng = np.random.default_rng(5)
    s, rot, ctr = 0.11 / 0.04, 0.7, np.array([2010.3, 1977.1])
    cs, sn = s * np.cos(rot), s * np.sin(rot)

    def outpix(inxys):
        d = np.asarray(inxys, dtype=np.float64) - ctr
        return np.stack([cs * d[:, 0] - sn * d[:, 1] + 79.5, sn * d[:, 0] + cs * d[:, 1] + 1e-5 * d[:, 0] ** 2 + 79.5], axis=1)

    ns = cfg.n1P + 2
    use = np.ones((ns, ns), dtype=bool)
    use[1, 2] = False
    mask = rng.random((4088, 4088)) < 0.9
    dp = DevicePartition(cfg, use)
    got = to_host(dp.partition(outpix, mask))
    want = OP.partition_pixels(outpix, mask, cfg, use)
    assert want["pix_count"].sum() > 5000 and got["n_cells"] >= 9
    for k in ("pix_count", "y_idx", "x_idx", "y_val", "x_val"):
        assert np.array_equal(got[k], want[k]), k
    small = DevicePartition(cfg, use, relax_coef=0.5)
    with pytest.raises(IndexError):
        small.partition(outpix, mask)


def test_device_partition_feeds_block():
    """f2 -> (a): InStamps binned on the device go into GpuBlock without a host round trip of the layers
    (partition.PartitionedBlock).  The block maps are bit-identical to those of the host path fed with the downloaded
    copies of the same lists, and the per-stamp pixel sets are those of the synthetic block's own binning."""
    import copy
    import types

    from pyimcom_b200.partition import DevicePartition, PartitionedBlock, to_host
    from pyimcom_b200.synth import SynthInStamp

    spec = cases.BLOCK_CASES["pad4"]
    blk = cases.make_block(spec)
    cfg = blk.cfg
    ns = cfg.n1P + 2
    sca, shift = 256, 100.0  # the synthetic detector coordinates are centred near 0: shift them into [0, sca)
    rng = np.random.default_rng(3)
    dp = DevicePartition(cfg, np.ones((ns, ns), dtype=bool), sca_nside=sca, sp_res=16)
    parts, twins = [], []
    for im in blk.inimages:
        vv, uu = np.mgrid[0:sca, 0:sca]
        indata = rng.standard_normal((cfg.n_inframe, sca, sca)).astype(np.float32)
        sin = im.outpix2world2inpix(np.asarray(blk.star_xy, dtype=np.float64)[None, :])[0]
        indata[0] = im.psf(uu - shift - sin[0], vv - shift - sin[1]).astype(np.float32)  # a star through this PSF
        part = dp.partition(lambda xy, im=im: im.inpix2outpix(np.asarray(xy, dtype=np.float64) - shift),
                            np.ones((sca, sca), dtype=bool), indata)
        assert part["is_relevant"]
        parts.append(part)
        h = to_host(part)
        tw = copy.copy(im)
        tw.pix_count, tw.x_val, tw.y_val, tw.data, tw.max_count = h["pix_count"], h["x_val"], h["y_val"], h["data"], h["max_count"]
        twins.append(tw)
        for j in range(ns):  # same pixel sets as the synthetic block's own binning (list order differs: cell-major)
            for i in range(ns):
                n = int(im.pix_count[j, i])
                assert n == int(h["pix_count"][j, i])
                assert np.array_equal(np.sort(im.x_val[j, i, :n]), np.sort(h["x_val"][j, i, :n]))
    pblk = PartitionedBlock(cfg, blk.inimages, parts, outwcs=blk.outwcs)
    hblk = types.SimpleNamespace(cfg=cfg, inimages=twins, n_inimage=len(twins), outwcs=blk.outwcs, this_sub=0,
                                 stamp_order=pblk.stamp_order)
    hblk.instamps = [[SynthInStamp(hblk, j, i) for i in range(ns)] for j in range(ns)]
    for j in range(ns):
        for i in range(ns):
            assert np.array_equal(pblk.instamps[j][i].x_val, hblk.instamps[j][i].x_val)
            assert np.array_equal(pblk.instamps[j][i].pix_cumsum, hblk.instamps[j][i].pix_cumsum)
    gd = GpuBlock(pblk, PSFTables(pblk, G.iD5512C, G.gridD5512C)).prepare().run()
    gh = GpuBlock(hblk, PSFTables(hblk, G.iD5512C, G.gridD5512C)).prepare().run()
    assert gd.h2d_bytes < gh.h2d_bytes
    md, mh = gd.download(), gh.download()
    for k in md:
        assert np.array_equal(md[k], mh[k]), k
    assert np.abs(md["out_map"][0, 0]).max() > 0.01  # the star is there


def test_sysmat_seam():
    """System-matrix seam: pyimcom_b200.sysmat.SysMatA / SysMatB driven exactly as the reference's OutStamp
    drives psfutil.SysMatA / SysMatB -- counting pass with sim_mode=True, caches cleared, then the assembly of
    coadd.py:1028-1082 -- give the oracle's A and -B/2 (P-f64), and the reference counts drain to zero."""
    from itertools import combinations

    from pyimcom_b200.sysmat import SysMatA, SysMatB

    spec = cases.BLOCK_CASES["pad4"]
    blk = cases.make_block(spec)
    cfg = blk.cfg
    otab = PSFTables(blk, R.iD5512C, R.gridD5512C)
    ns = cfg.n1P + 2
    stamps = list(spec["stamps"])
    blk.outstamps = [[None] * ns for _ in range(ns)]
    for (j, i) in stamps:
        blk.outstamps[j][i] = OracleOutStamp(blk, otab, j, i)
    sa, sb = SysMatA(blk), SysMatB(blk)
    for (j, i) in stamps:  # OutStamp.__init__ in sim mode (coadd.py:860-867)
        o = blk.outstamps[j][i]
        for ji in o.ji_st_in_s:
            assert sa.get_iisubmat(ji, ji, sim_mode=True) is None
            assert sb.get_iosubmat(ji, (j, i), sim_mode=True) is None
        for pair in combinations(o.ji_st_in_s, 2):
            sa.get_iisubmat(*pair, sim_mode=True)
    assert sa.iisubmats and all(v is None for v in sa.iisubmats.values())
    sa.iisubmats.clear()  # coadd.py:2062-2063
    sb.iopsfovls.clear()
    with pytest.raises(AssertionError):
        sa.get_iisubmat((2, 2), (1, 1))
    with pytest.raises(AssertionError):
        sb.get_iosubmat((0, 0), (2, 2))
    for (j, i) in stamps:  # OutStamp._build_system_matrices (coadd.py:1028-1082) through the seam
        o = blk.outstamps[j][i]
        cs = o.inpix_cumsum.astype(np.int64)
        n = int(cs[-1])
        sysmata = np.zeros((n, n))
        for idx, ji, sel in zip(range(9), o.ji_st_in_s, o.selections):
            sub = sa.get_iisubmat(ji, ji)
            assert sub.dtype == np.float64 and sub.shape[0] == sub.shape[1] == int(blk.instamps[ji[0]][ji[1]].pix_cumsum[-1])
            if sel is not None:
                sub = sub[np.ix_(sel, sel)]
            sysmata[cs[idx]:cs[idx + 1], cs[idx]:cs[idx + 1]] = sub
        for idx_s, pair, sels in zip(combinations(range(9), 2), combinations(o.ji_st_in_s, 2),
                                     combinations(o.selections, 2)):
            sub = sa.get_iisubmat(*pair)
            if sels[0] is not None:
                sub = sub[np.ix_(sels[0], sels[1])] if sels[1] is not None else sub[sels[0], :]
            elif sels[1] is not None:
                sub = sub[:, sels[1]]
            r0, r1, c0, c1 = cs[idx_s[0]], cs[idx_s[0] + 1], cs[idx_s[1]], cs[idx_s[1] + 1]
            sysmata[r0:r1, c0:c1] = sub
            sysmata[c0:c1, r0:r1] = sub.T
        mhalfb = np.zeros((cfg.n_out, cfg.n2f**2, n))
        for idx, ji in zip(range(9), o.ji_st_in_s):
            mhalfb[:, :, cs[idx]:cs[idx + 1]] = sb.get_iosubmat(ji, (j, i))
        o.build_system_matrices()
        assert rel(sysmata, o.sysmata) < P64 and rel(mhalfb, o.mhalfb) < P64
        assert np.abs(sysmata).max() > 0 and np.abs(mhalfb).max() > 0
    assert not sa.iisubmats and not sb.iopsfovls
    assert not sa.iisubmats_ref.any() and not sb.iopsfovls_ref.any()
