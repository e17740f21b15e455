"""GPU parity of the Gram / distance kernels, called through the C ABI (ctypes binding).

Truth is the oracle's float64 direct differencing on the float32 inputs; the tolerance is
north_star's rel 1e-5 / abs 1e-6.  The reference-order float32 oracle (ref32) is compared
wherever it is itself within tolerance of truth (it is not on self-Gram diagonals, p < 2).
"""
import numpy as np
import pytest
import torch

from jaxrk_b200 import _native as nat
from oracle import jaxrk_oracle as orc
from gpu_util import ATOL, RTOL, assert_close_truth, cu, f32, gauss, kparams

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kname", ["sqexp", "laplace", "gg15", "gg05"])
@pytest.mark.parametrize("path", [nat.PATH_DIRECT, nat.PATH_AUTO])
def test_cfg1_gram_1024_d8(kname, path):
    """BASELINE config 1: N = M = 1024, d = 8 (full output checked)."""
    X, Y = gauss(0, 1024, 8), gauss(1, 1024, 8)
    s, scale, p, nconst = kparams(kname, np.sqrt(8.0))
    K = nat.gram(cu(X), cu(Y), scale, p, nconst, path=path).cpu().numpy()
    truth = orc.gengauss_gram_truth64(X, Y, s, p)
    assert_close_truth(K, truth, kname)
    ref = orc.gengauss_gram_ref32(X, Y, s, p)
    ok = orc.within_tol(ref, truth)
    assert ok.mean() > 0.999
    assert (np.abs(K - ref)[ok] <= 2 * (ATOL + RTOL * np.abs(ref[ok]))).all()


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 7, 8, 9, 12, 13, 16, 17, 24, 31, 32, 33, 48, 63])
def test_direct_path_all_small_d_ragged(d):
    N, M = 257, 131  # neither a multiple of the 128 x 128 tile
    X, Y = gauss(d, N, d), gauss(100 + d, M, d)
    for kname in ("sqexp", "laplace", "gg15"):
        s, scale, p, nconst = kparams(kname, np.sqrt(d) * 1.3)
        K = nat.gram(cu(X), cu(Y), scale, p, nconst, path=nat.PATH_DIRECT).cpu().numpy()
        assert K.shape == (N, M)
        assert_close_truth(K, orc.gengauss_gram_truth64(X, Y, s, p), f"{kname} d={d}")


@pytest.mark.parametrize("d", [64, 65, 96, 100, 128, 200, 256])
@pytest.mark.parametrize("path", [nat.PATH_DIRECT, nat.PATH_TF32X3, nat.PATH_AUTO])
def test_large_d_both_paths(d, path):
    N, M = 300, 200
    X, Y = gauss(d, N, d), gauss(100 + d, M, d)
    for kname in ("sqexp", "laplace", "gg15"):
        s, scale, p, nconst = kparams(kname, np.sqrt(d))
        K = nat.gram(cu(X), cu(Y), scale, p, nconst, path=path).cpu().numpy()
        assert_close_truth(K, orc.gengauss_gram_truth64(X, Y, s, p), f"{kname} d={d} path={path}")


@pytest.mark.parametrize("N,M", [(1, 1), (1, 300), (300, 1), (2, 3), (127, 129), (128, 128), (513, 4)])
def test_edge_shapes(N, M):
    X, Y = gauss(N, N, 6), gauss(M + 7, M, 6)
    s, scale, p, nconst = kparams("gg15", 2.0)
    K = nat.gram(cu(X), cu(Y), scale, p, nconst).cpu().numpy()
    assert K.shape == (N, M)
    assert_close_truth(K, orc.gengauss_gram_truth64(X, Y, s, p))


def test_empty_inputs():
    Y = cu(gauss(0, 5, 3))
    E = torch.empty((0, 3), dtype=torch.float32, device=Y.device)
    assert tuple(nat.gram(E, Y, 1.0, 2.0, 1.0).shape) == (0, 5)
    assert tuple(nat.gram(Y, E, 1.0, 2.0, 1.0).shape) == (5, 0)


@pytest.mark.parametrize("d,path", [(5, nat.PATH_DIRECT), (16, nat.PATH_DIRECT), (64, nat.PATH_TF32X3), (96, nat.PATH_TF32X3)])
@pytest.mark.parametrize("kname", ["sqexp", "laplace", "gg15", "gg05"])
def test_self_gram_is_bitwise_symmetric_with_exact_diagonal(d, path, kname):
    """jaxrk/utilities/gram.py:18 asserts G == G.T; direct differencing gives the exact
    nconst on the diagonal where the reference's expansion form is off by up to 1e-4."""
    X = gauss(3, 777, d)
    s, scale, p, nconst = kparams(kname, np.sqrt(d))
    Xd = cu(X)
    truth = orc.gengauss_gram_truth64(X, None, s, p)
    for Y in (None, Xd):  # Y = None and Y aliasing X: the C ABI's definition of a self-Gram
        K = nat.gram(Xd, Y, scale, p, nconst, path=path).cpu().numpy()
        assert (K == K.T).all()
        np.testing.assert_allclose(np.diag(K), nconst, rtol=1e-6 if path == nat.PATH_DIRECT else 1e-5)
        assert_close_truth(K, truth, kname)
    # an equal COPY of X is not a self-Gram for the library: same tolerance, symmetry only to rounding
    K = nat.gram(Xd, cu(X), scale, p, nconst, path=path).cpu().numpy()
    assert_close_truth(K, truth, kname)
    np.testing.assert_allclose(K, K.T, rtol=2e-6, atol=1e-9)
    if path == nat.PATH_DIRECT:
        assert (K == K.T).all()


def test_near_duplicate_points_tight_at_zero_distance():
    """north_star: tolerance 'tightest at near-zero distances'."""
    rng = np.random.RandomState(9)
    X = rng.randn(200, 16).astype(f32) * 3
    Y = (X + rng.randn(200, 16).astype(f32) * 1e-4).astype(f32)
    for kname in ("laplace", "gg05", "sqexp"):
        s, scale, p, nconst = kparams(kname, 1.0)
        K = nat.gram(cu(X), cu(Y), scale, p, nconst, path=nat.PATH_DIRECT).cpu().numpy()
        assert_close_truth(K, orc.gengauss_gram_truth64(X, Y, s, p), kname)


def test_strided_and_unaligned_operands():
    base = cu(gauss(0, 300, 24))
    X = base[:, 3:16]            # ld 24, offset 3 floats: not 16-byte aligned, d = 13
    Y = base[7:150, 5:18]
    s, scale, p, nconst = kparams("laplace", 3.0)
    out = torch.zeros((300, 160), dtype=torch.float32, device=base.device)
    Kv = out[:, 1:144]           # unaligned output view with ld 160
    nat.gram(X, Y, scale, p, nconst, out=Kv)
    truth = orc.gengauss_gram_truth64(X.cpu().numpy(), Y.cpu().numpy(), s, p)
    assert_close_truth(Kv.cpu().numpy(), truth)
    assert (out[:, 0] == 0).all() and (out[:, 144:] == 0).all(), "wrote outside the output view"


def test_per_dimension_scale_and_distance_output(golden):
    X, Y, ls = golden["gram_X"], golden["gram_Y"], golden["perdim_ls"]
    ds = (f32(1) / ls).astype(f32)
    D = nat.dist(cu(X), cu(Y), 1.0, 1.5, dim_scale=ds).cpu().numpy()
    np.testing.assert_allclose(D, golden["perdim_dist_XY"], rtol=2e-5, atol=2e-6)
    a, b = golden["eu_a"], golden["eu_b"]
    for power in (1, 2):
        got = nat.dist(cu(a), cu(b), 1.0, float(power)).cpu().numpy()
        np.testing.assert_allclose(got, golden[f"eu_simple_p{power}_ab"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(got, np.sqrt(orc.sqeucldist_truth64(a, b)) ** power, rtol=1e-6, atol=1e-7)
    self_d = nat.dist(cu(a), None, 1.0, 1.0).cpu().numpy()
    assert (np.diag(self_d) == 0).all() and (self_d == self_d.T).all()


@pytest.mark.parametrize("name,mk,p", [("gauss", lambda: orc.simple_scaler_s(orc.gauss_length_scale(1.3)), 2.0),
                                       ("laplace", lambda: orc.simple_scaler_s(0.7), 1.0),
                                       ("gg15", lambda: orc.simple_scaler_s(2.0), 1.5),
                                       ("gg05", lambda: orc.simple_scaler_s(0.9), 0.5)])
def test_golden_grams_from_the_reference_code(golden, name, mk, p):
    s = mk()
    nconst = float(golden[f"gram_{name}_nconst"].reshape(-1)[0])
    X, Y = golden["gram_X"], golden["gram_Y"]
    K = nat.gram(cu(X), cu(Y), float(s[0, 0]), p, nconst).cpu().numpy()
    np.testing.assert_allclose(K, golden[f"gram_{name}_XY"], rtol=2e-5, atol=1e-6)
    Ks = nat.gram(cu(X), None, float(s[0, 0]), p, nconst).cpu().numpy()
    off = ~np.eye(len(X), dtype=bool)
    np.testing.assert_allclose(Ks[off], golden[f"gram_{name}_XX"][off], rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("d", [16, 24, 32, 48, 63])
def test_tensor_path_for_small_d(d):
    """PATH_TF32X3 pads d < 64 to 64 columns; AUTO picks it for 16 <= d < 64 once N*M >= 2^24."""
    N, M = 4200, 4100   # N*M >= 2^24: AUTO takes the tensor path
    X, Y = gauss(d, N, d), gauss(100 + d, M, d)
    for kname in ("sqexp", "laplace", "gg15"):
        s, scale, p, nconst = kparams(kname, np.sqrt(d))
        Ka = nat.gram(cu(X), cu(Y), scale, p, nconst, path=nat.PATH_AUTO)
        Kt = nat.gram(cu(X), cu(Y), scale, p, nconst, path=nat.PATH_TF32X3)
        assert torch.equal(Ka, Kt)
        rows = np.arange(0, N, 41)
        truth = orc.gengauss_gram_truth64(X[rows], Y, s, p)
        assert_close_truth(Kt.cpu().numpy()[rows], truth, f"{kname} d={d}")
        Kd = nat.gram(cu(X), cu(Y), scale, p, nconst, path=nat.PATH_DIRECT)
        assert torch.allclose(Kd, Kt, rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("d,path", [(16, nat.PATH_DIRECT), (16, nat.PATH_AUTO), (64, nat.PATH_TF32X3)])
def test_full_size_properties(d, path):
    """At sizes the oracle cannot hold: row-block independence, symmetry, sampled entries."""
    N = 16384
    X = gauss(0, N, d)
    s, scale, p, nconst = kparams("sqexp", np.sqrt(d))
    Xd = cu(X)
    K = nat.gram(Xd, None, scale, p, nconst, path=path)
    assert torch.equal(K, K.T)
    blk = nat.gram(Xd[4096:4096 + 512], Xd, scale, p, nconst, path=path)
    # (under AUTO the 512-row block is small enough for the direct path while the full Gram takes the tensor path)
    tight = path != nat.PATH_AUTO
    assert torch.allclose(K[4096:4096 + 512], blk, rtol=1e-6 if tight else 2e-5, atol=1e-9 if tight else 2e-6)
    rows = np.random.RandomState(1).choice(N, 64, replace=False)
    truth = orc.gengauss_gram_truth64(X[rows], X, s, p)
    assert_close_truth(K[torch.from_numpy(rows).to(K.device)].cpu().numpy(), truth)


def test_host_buffer_entry_point_streams_row_blocks():
    N, M, d = 3000, 2500, 16
    X, Y = gauss(0, N, d), gauss(1, M, d)
    s, scale, p, nconst = kparams("gg15", 4.0)
    K = np.empty((N, M), f32)
    nat.gram_host(X, Y, K, scale, p, nconst)
    assert_close_truth(K, orc.gengauss_gram_truth64(X, Y, s, p))
    Ks = np.empty((N, N), f32)
    nat.gram_host(X, None, Ks, scale, p, nconst)
    assert (Ks == Ks.T).all()


def test_invalid_arguments_raise():
    X = cu(gauss(0, 8, 3))
    with pytest.raises(nat.JrkError) as e:
        nat.gram(X, None, 1.0, 2.5, 1.0)
    assert e.value.code == 1
    with pytest.raises(nat.JrkError):
        nat.gram(X, None, -1.0, 2.0, 1.0)
    with pytest.raises(ValueError):
        nat.gram(X, cu(gauss(1, 8, 4)), 1.0, 2.0, 1.0)
    with pytest.raises(TypeError):
        nat.gram(X.double(), None, 1.0, 2.0, 1.0)


@pytest.mark.parametrize("d", [20, 33, 64, 100, 128])
@pytest.mark.parametrize("case", ["huge", "tiny", "mixed_columns", "zero_rows", "offset"])
def test_fp16_operand_path_is_robust_to_input_magnitudes(d, case):
    """32 < d <= 128 runs on fp16 hi / lo operands normalised by powers of two taken from the data (gram_tf32_kernel,
    H16): magnitudes far outside fp16's range, columns of very different scale, zero rows and a large common offset
    must all stay within tolerance of the float64 truth."""
    rng = np.random.RandomState(d)
    N, M = 260, 190
    X, Y = rng.randn(N, d), rng.randn(M, d)
    ls = np.sqrt(d)
    if case == "huge":
        X, Y, ls = X * 3e4, Y * 3e4, ls * 3e4
    elif case == "tiny":
        X, Y, ls = X * 2e-6, Y * 2e-6, ls * 2e-6
    elif case == "mixed_columns":
        col = np.where(np.arange(d) % 7 == 0, 50.0, 1e-3)
        X, Y, ls = X * col, Y * col, 50.0 * np.sqrt(d / 7)
    elif case == "zero_rows":
        X[::3] = 0.0
        Y[5] = 0.0
    elif case == "offset":
        X, Y = X + 40.0, Y + 40.0          # |x|^2 >> |x - y|^2: the expansion form is at its worst
    X, Y = X.astype(f32), Y.astype(f32)
    for kname in ("sqexp", "laplace"):
        s, scale, p, nconst = kparams(kname, ls)
        K = nat.gram(cu(X), cu(Y), scale, p, nconst, path=nat.PATH_TF32X3).cpu().numpy()
        assert_close_truth(K, orc.gengauss_gram_truth64(X, Y, s, p), f"{case} {kname} d={d}")
        Ks = nat.gram(cu(X), None, scale, p, nconst, path=nat.PATH_TF32X3).cpu().numpy()
        assert (Ks == Ks.T).all()
        assert_close_truth(Ks, orc.gengauss_gram_truth64(X, X, s, p), f"{case} {kname} d={d} self")
