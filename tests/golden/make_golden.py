"""Generate tests/golden/*.npz by running the REFERENCE'S OWN PYTHON CODE.

Run in the build container only (needs /root/reference; never at test time):

    python tests/golden/make_golden.py

jax / jaxlib / flax are absent from this image (and reference HEAD imports symbols
that no released jax still has), so the reference package is imported over a
numpy-backed stand-in for the `jax` namespace defined below.  What this pins is the
reference's *Python-level op sequence* (which jnp calls, in which order, on which
operands, with float32 storage); what it cannot pin is XLA's own rounding inside
exp/pow/GEMM, which is covered to ~1e-7 by the authors' printed notebook outputs
(golden G1 in tests/test_oracle.py).

Stand-in rules (to mimic jax without x64):
  * array creation defaults to float32 (jnp.array/ones/zeros/eye/linspace/atleast_*),
  * jnp.float64 is float32, and `.astype(float)` inside jaxrk.rkhs.vector resolves
    to float32 (module-global `float` injected after import; quirk Q4 of SURVEY.md),
  * jax.scipy.special.gammaln computes in float32,
  * jit is the identity; grad/vmap raise if reached (not on the hot path).
All test inputs are cast to float32 before they enter the reference, which is what
jax does on entry (ints and float64 are truncated to float32).
"""
import os
import sys
import types
import warnings

import numpy as onp
import scipy.special
import scipy.stats
import scipy.linalg

warnings.filterwarnings("ignore")
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
F32 = onp.float32


# ----------------------------------------------------------------------------------
# the numpy-backed `jax` stand-in
# ----------------------------------------------------------------------------------
def _f32ify(a):
    a = onp.asarray(a)
    if a.dtype == onp.float64:
        return a.astype(F32)
    return a


class _JnpModule(types.ModuleType):
    float64 = F32
    float32 = F32
    ndarray = onp.ndarray
    newaxis = None

    def __getattr__(self, name):
        return getattr(onp, name)

    @staticmethod
    def array(obj, dtype=None, **kw):
        return _f32ify(onp.array(obj, dtype=dtype, **kw)) if dtype is None else onp.array(obj, dtype=dtype, **kw)

    @staticmethod
    def asarray(obj, dtype=None):
        return _f32ify(onp.asarray(obj)) if dtype is None else onp.asarray(obj, dtype=dtype)

    @staticmethod
    def ones(shape, dtype=None):
        return onp.ones(shape, dtype=F32 if dtype is None else dtype)

    @staticmethod
    def zeros(shape, dtype=None):
        return onp.zeros(shape, dtype=F32 if dtype is None else dtype)

    @staticmethod
    def eye(n, m=None, dtype=None):
        return onp.eye(n, m, dtype=F32 if dtype is None else dtype)

    @staticmethod
    def linspace(*a, **kw):
        return onp.linspace(*a, **kw).astype(F32)

    @staticmethod
    def clip(a, a_min=None, a_max=None):
        return onp.clip(a, a_min, a_max)

    @staticmethod
    def atleast_1d(x):
        return _f32ify(onp.atleast_1d(x))

    @staticmethod
    def atleast_2d(x):
        return _f32ify(onp.atleast_2d(x))


def _unavailable(*a, **k):
    raise RuntimeError("not available in the numpy stand-in (off the hot path)")


def install_shim():
    jnp = _JnpModule("jax.numpy")
    jnp.linalg = onp.linalg

    jax = types.ModuleType("jax")
    jax.numpy = jnp
    jax.jit = lambda f=None, **kw: f if f is not None else (lambda g: g)
    jax.grad = lambda f, *a, **k: _unavailable
    jax.vmap = lambda f, *a, **k: _unavailable

    jsp = types.ModuleType("jax.scipy")
    special = types.ModuleType("jax.scipy.special")
    special.gammaln = lambda x: scipy.special.gammaln(onp.asarray(x, dtype=F32)).astype(F32)[()]
    special.logsumexp = scipy.special.logsumexp
    special.expit = scipy.special.expit
    special.logit = scipy.special.logit
    jsp.special = special
    jsp.stats = scipy.stats
    jsp.linalg = scipy.linalg

    random = types.ModuleType("jax.random")
    random.PRNGKey = lambda seed: onp.random.RandomState(seed)
    random.randint = _unavailable
    random.split = _unavailable

    lax = types.ModuleType("jax.lax")
    lax.scan = _unavailable
    lax.map = _unavailable

    ops = types.ModuleType("jax.ops")
    ops.index_update = _unavailable

    interp = types.ModuleType("jax.interpreters")
    xla = types.ModuleType("jax.interpreters.xla")
    xla.DeviceArray = onp.ndarray
    interp.xla = xla

    jax.scipy, jax.random, jax.lax, jax.ops, jax.interpreters = jsp, random, lax, ops, interp

    flax = types.ModuleType("flax")
    linen = types.ModuleType("flax.linen")
    linen.Module = object
    linen_module = types.ModuleType("flax.linen.module")
    linen_module.compact = lambda f: f
    linen.module = linen_module
    flax.linen = linen

    pathlib2 = types.ModuleType("pathlib2")
    import pathlib
    pathlib2.Path = pathlib.Path

    mods = {
        "jax": jax, "jax.numpy": jnp, "jax.scipy": jsp, "jax.scipy.special": special,
        "jax.scipy.stats": scipy.stats, "jax.scipy.linalg": scipy.linalg, "jax.random": random,
        "jax.lax": lax, "jax.ops": ops, "jax.interpreters": interp, "jax.interpreters.xla": xla,
        "flax": flax, "flax.linen": linen, "flax.linen.module": linen_module, "pathlib2": pathlib2,
    }
    sys.modules.update(mods)


def import_reference():
    install_shim()
    sys.path.insert(0, REF)
    import jaxrk  # noqa: F401  (the unmodified reference package)
    import jaxrk.rkhs.vector as vector
    vector.float = F32  # `.astype(float)` -> float32, as under jax without x64
    return jaxrk


# ----------------------------------------------------------------------------------
# cases
# ----------------------------------------------------------------------------------
def main():
    jaxrk = import_reference()
    from jaxrk.kern import GenGaussKernel
    from jaxrk.reduce import Prefactors, BalancedRed, LinearReduce, Sum, Mean, Center, TileView
    from jaxrk.rkhs import FiniteVec, inner, FiniteOp, CovOp, Cov_inv, Cov_solve, Cmo, Cdo
    from jaxrk.utilities.eucldist import eucldist
    from jaxrk.utilities.distances import dist

    out = {}
    rng = onp.random.RandomState(20261019)

    # ---- A. tests/test_kern.py::test_rbf shaped known answers ---------------------
    x8 = onp.arange(8).reshape(-1, 1).astype(F32)
    ls_grid = onp.linspace(0.9, 2.0, 5).astype(F32)
    sh_grid = onp.linspace(0.9, 2.0, 5).astype(F32)
    rows, nconsts, variances, params = [], [], [], []
    for ls in ls_grid:
        kernels = [("gauss", GenGaussKernel.make_gauss(length_scale=ls), 2.0),
                   ("laplace", GenGaussKernel.make_laplace(length_scale=ls), 1.0)]
        for sh in sh_grid:
            kernels.append(("gengauss", GenGaussKernel.make(shape=float(sh), length_scale=ls), float(sh)))
        for name, k, sh in kernels:
            g = k(x8)
            rows.append(onp.asarray(g[0, :], dtype=F32))
            nconsts.append(F32(onp.asarray(k.nconst).reshape(-1)[0]))
            variances.append(F32(onp.asarray(k.var()).reshape(-1)[0]))
            params.append((("gauss", "laplace", "gengauss").index(name), float(ls), sh))
    out["rbf_row0"] = onp.stack(rows)
    out["rbf_nconst"] = onp.array(nconsts, F32)
    out["rbf_var"] = onp.array(variances, F32)
    out["rbf_params"] = onp.array(params, onp.float64)  # (kind, length_scale, shape)

    # ---- B. full Gram matrices, X != Y and self ------------------------------------
    X = rng.randn(37, 5).astype(F32)
    Y = (rng.randn(29, 5) * 1.5 + 0.3).astype(F32)
    out["gram_X"], out["gram_Y"] = X, Y
    kdefs = {"gauss": GenGaussKernel.make_gauss(1.3), "laplace": GenGaussKernel.make_laplace(0.7),
             "gg15": GenGaussKernel.make(2.0, 1.5), "gg05": GenGaussKernel.make(0.9, 0.5)}
    for name, k in kdefs.items():
        out[f"gram_{name}_XY"] = k(X, Y)
        out[f"gram_{name}_XX"] = k(X)
        out[f"gram_{name}_diag"] = k(X, X[::-1].copy(), diag=True)
        out[f"gram_{name}_nconst"] = onp.asarray(k.nconst)
    # per-dimension length scales: only the distance is usable (quirk Q2)
    ls_vec = onp.array([0.5, 1.0, 1.5, 2.0, 2.5], F32)
    kd = GenGaussKernel.make(ls_vec, 1.5)
    out["perdim_ls"] = ls_vec
    out["perdim_dist_XY"] = kd.dist(X, Y)
    out["perdim_nconst"] = onp.asarray(kd.nconst)

    # ---- C. eucldist / dist (tests/test_utilities.py) ------------------------------
    a = rng.randn(20, 3).astype(F32)
    b = rng.randn(30, 3).astype(F32)
    out["eu_a"], out["eu_b"] = a, b
    for variant in ("simple", "extension"):
        for power in (1, 2):
            out[f"eu_{variant}_p{power}_ab"] = eucldist(a, b, power=power, variant=variant)
            out[f"eu_{variant}_p{power}_aa"] = eucldist(a, a, power=power, variant=variant)
    out["dist_p1_ab"] = dist(a, b, power=1)
    out["dist_p2_ab"] = dist(a, b, power=2)

    # ---- D. FiniteVec.inner with reduce chains -------------------------------------
    # D1: examples/Quick_start.ipynb cells 3-5 (golden G1 re-derived through the shim)
    k = GenGaussKernel.make_laplace(1.0)
    pts = onp.arange(6).reshape(-1, 1).astype(F32)
    v1 = FiniteVec(k, pts, [Prefactors(onp.ones(6, F32))])
    w1 = FiniteVec(k, pts, [Prefactors(onp.ones(6, F32) / 6), BalancedRed(3)])
    e1 = FiniteVec.construct_RKHS_Elem(k, pts, onp.ones(6, F32))
    out["qs_inner_v1_w1"] = inner(v1, w1)
    out["qs_inner_v1_e1"] = inner(v1, e1)

    # D2: a battery of chains on random points
    kg = GenGaussKernel.make(1.7, 1.5)
    P = rng.randn(24, 4).astype(F32)
    Q = rng.randn(36, 4).astype(F32)
    pp = rng.rand(24).astype(F32)
    pq = rng.rand(36).astype(F32)
    A_p = rng.randn(3, 24).astype(F32)
    A_q = rng.randn(5, 36).astype(F32)
    A_q12 = rng.randn(2, 12).astype(F32)
    out.update(ch_P=P, ch_Q=Q, ch_pp=pp, ch_pq=pq, ch_Ap=A_p, ch_Aq=A_q, ch_Aq12=A_q12)
    chains = {
        "none": [],
        "pref": [Prefactors(pp)],
        "sum": [Sum()],
        "mean": [Mean()],
        "pref_sum": [Prefactors(pp), Sum()],
        "bal4": [BalancedRed(4)],
        "bal4avg": [BalancedRed(4, True)],
        "pref_bal3": [Prefactors(pp), BalancedRed(3)],
        "lin": [LinearReduce(A_p)],
        "pref_lin": [Prefactors(pp), LinearReduce(A_p)],
        "center": [Center()],
        "bal2_center": [BalancedRed(2), Center()],
    }
    chains_q = {
        "none": [],
        "pref": [Prefactors(pq)],
        "sum": [Sum()],
        "mean": [Mean()],
        "lin": [LinearReduce(A_q)],
        "pref_lin": [Prefactors(pq), LinearReduce(A_q)],
        "bal3_lin": [BalancedRed(3), LinearReduce(A_q12)],
        "bal6avg": [BalancedRed(6, True)],
        "tile": [Mean(), TileView(4)],
    }
    for nx, cx in chains.items():
        for ny, cy in chains_q.items():
            out[f"ch_inner__{nx}__{ny}"] = inner(FiniteVec(kg, P, cx), FiniteVec(kg, Q, cy))
    for nx, cx in chains.items():
        out[f"ch_self__{nx}"] = FiniteVec(kg, P, cx).inner()
    # evaluation of an RKHS element at points: vec(pts)  (rkhs/vector.py:203-204)
    el = FiniteVec.construct_RKHS_Elem(kg, P, pp)
    out["ch_eval_elem_at_Q"] = el(Q)
    out["ch_eval_lin_at_Q"] = FiniteVec(kg, P, [LinearReduce(A_p)])(Q)

    # ---- E. operators (rkhs/operator.py) -------------------------------------------
    kx = GenGaussKernel.make_gauss(0.9)
    ky = GenGaussKernel.make_laplace(1.1)
    n = 16
    Xo = rng.randn(n, 2).astype(F32)
    Yo = (Xo[:, :1] * 0.7 + 0.3 * rng.randn(n, 1)).astype(F32)
    Xt = rng.randn(7, 2).astype(F32)
    Yt = onp.linspace(-2, 2, 9).astype(F32)[:, None]
    Rf = rng.randn(12, 1).astype(F32)
    M_op = rng.randn(n, n).astype(F32)
    out.update(op_X=Xo, op_Y=Yo, op_Xt=Xt, op_Yt=Yt, op_R=Rf, op_M=M_op)
    inp, outp, ref = FiniteVec(kx, Xo), FiniteVec(ky, Yo), FiniteVec(ky, Rf)
    op = FiniteOp(inp, outp, M_op)
    res = op @ FiniteVec(kx, Xt)
    out["op_matmul_vec_lin"] = res.reduce[-1].linear_map          # (7 x 16)
    out["op_matmul_vec_eval"] = res(Yt)                           # (7 x 9)
    res1 = op @ FiniteVec(kx, Xt[:1])
    out["op_matmul_vec1_eval"] = res1(Yt)                         # LinearReduce + Sum
    out["op_matmul_arr_eval"] = (op @ Xt)(Yt)
    op2 = FiniteOp(outp, inp, M_op.T.copy())
    out["op_matmul_op_matr"] = (op @ op2).matr
    cov = CovOp(inp)
    out["cov_inv_matr"] = Cov_inv(cov, regul=0.1).matr
    cmo = Cmo(inp, outp, regul=0.1)
    out["cmo_matr"] = cmo.matr
    out["cmo_inp_lin"] = cmo.inp_feat.reduce[-1].linear_map
    out["cmo_apply_eval"] = (cmo @ FiniteVec(kx, Xt))(Yt)
    cdo = Cdo(inp, outp, ref, regul=0.1)
    out["cdo_matr"] = cdo.matr
    out["cdo_apply_eval"] = (cdo @ FiniteVec(kx, Xt))(Yt)

    path = os.path.join(HERE, "reference_numpy_shim.npz")
    onp.savez_compressed(path, **{k_: onp.asarray(v) for k_, v in out.items()})
    print("wrote", path, len(out), "arrays,", os.path.getsize(path), "bytes")
    dts = {str(onp.asarray(v).dtype) for v in out.values()}
    print("dtypes:", dts)


if __name__ == "__main__":
    main()
