"""RKHS distances and representer selection from Gram matrices -- drop-in for jaxrk/utilities/gram.py:8-46, 51-53.

These are consumers of the hot path: the Gram matrices they take come from the fused kernels (`kernel(points)`,
`FiniteVec.inner`), what happens here is O(n^2) array arithmetic on those results.  The scipy-based positive projection
(`gram_projection(method="pos_proj")`, gram.py:54-66) is a host optimiser loop and out of scope.
"""
import torch

__all__ = ["rkhs_gram_cdist", "rkhs_gram_cdist_ignore_const", "rkhs_gram_cdist_unchecked", "choose_representer",
           "choose_representer_from_gram", "gram_projection"]


def _powered(sq, power: float):
    return sq if power == 2. else torch.pow(sq, power / 2.)


def rkhs_gram_cdist_unchecked(G_ab, G_a, G_b, power: float = 2.):
    """||a_i - b_j||_H^power = (G_a[i, i] + G_b[j, j] - 2 G_ab[i, j])^(power / 2)   (gram.py:30-35)."""
    return _powered(torch.diagonal(G_a)[:, None] + torch.diagonal(G_b)[None, :] - 2 * G_ab, power)


def rkhs_gram_cdist_ignore_const(G_ab, G_b, power: float = 2.):
    """The same without the term that is constant in j's counterpart: G_b[j, j] - 2 G_ab[i, j]   (gram.py:23-28)."""
    return _powered(torch.diagonal(G_b)[None, :] - 2 * G_ab, power)


def rkhs_gram_cdist(G_ab, G_a=None, G_b=None, power: float = 2.):
    """gram.py:8-21.  With a single Gram the reference asserts exact symmetry (`np.all(G_ab == G_ab.T)`); the
    self-Gram kernels of this package guarantee it bitwise."""
    assert G_ab.dim() == 2
    if G_a is not None:
        assert G_a.dim() == 2 and G_a.shape[0] == G_a.shape[1] and G_ab.shape[0] == G_a.shape[0]
    if G_b is not None:
        assert G_b.shape[0] == G_b.shape[1] and G_ab.shape[1] == G_b.shape[1]
    if G_a is None or G_b is None:
        assert G_a is None and G_b is None, 'Either none or both of G_repr, G_orig should be None'
        assert G_ab.shape[0] == G_ab.shape[1] and bool(torch.all(G_ab == G_ab.T))
        G_a = G_b = G_ab
    return rkhs_gram_cdist_unchecked(G_ab, G_a, G_b, power)


def choose_representer_from_gram(G, factors):
    """Index of the point whose feature is closest in RKHS norm to sum_i factors[i] phi(x_i)   (gram.py:41-46)."""
    factors = torch.as_tensor(factors, dtype=G.dtype, device=G.device).reshape(-1)
    fG = factors @ G
    d2 = (factors @ fG) + torch.diagonal(G) - 2 * fG
    return int(torch.argmin(d2))


def choose_representer(support_points, factors, kernel):
    """gram.py:38-39: the Gram of the support points comes from the fused self-Gram kernel."""
    return choose_representer_from_gram(kernel(support_points), factors)


def gram_projection(G_orig_repr, G_orig=None, G_repr=None, method: str = "representer"):
    """gram.py:51-53: for every original element the index of the nearest representer."""
    if method == "representer":
        return torch.argmin(rkhs_gram_cdist(G_orig_repr, G_repr, G_orig), 0)
    if method == "pos_proj":
        raise NotImplementedError("gram_projection(method='pos_proj') is a scipy L-BFGS loop over jax.grad in the reference "
                                  "(gram.py:54-66): out of scope of the accelerated path")
    assert False, "No valid method selected"
