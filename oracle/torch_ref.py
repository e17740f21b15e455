"""Reference op sequence on torch-CPU, eager, op by op.  TEST / BASELINE INFRASTRUCTURE ONLY
(see the header of oracle/jaxrk_oracle.py for who may import this).

The reference executes the Gram as one jitted expansion-form GEMM followed by ~9 separate eager
jax.numpy element-wise dispatches, each a multi-threaded, vectorised XLA-CPU
loop.  jax/jaxlib are absent from the image; torch-CPU eager float32 is the closest analogue
of that execution model (library GEMM + one vectorised multi-threaded pass per op), and it is
what bench.py's cpu_baseline / --impl reference legs time.  Row blocks bound the ~10 N x M
temporaries the reference would allocate.  Lines followed:
  jaxrk/utilities/eucldist.py:10-18,48   kern/util.py:115   kern/rbf.py:105
  reduce/lincomb.py:137-140 (A @ gram)   rkhs/vector.py:72-74
"""
import torch


def gram_block(X, Y, s: float, p: float, nconst: float) -> torch.Tensor:
    a2 = torch.einsum("ij,ij->i", X, X)                  # eucldist.py:12
    b2 = torch.einsum("ij,ij->i", Y, Y)                  # :14
    sq = a2[:, None] + b2 - 2 * X @ Y.T                  # :18
    r = torch.pow(torch.clip(sq, 0.), 1. / 2.)           # :48  (power = 1. from kern/util.py:115)
    t = s * r                                            # SimpleScaler (kern/util.py:62-69)
    u = t ** p                                           # kern/util.py:115
    return nconst * torch.exp(-u)                        # kern/rbf.py:105


def gram(X, Y, s, p, nconst, block=4096):
    out = torch.empty((X.shape[0], Y.shape[0]), dtype=torch.float32)
    for i in range(0, X.shape[0], block):
        out[i:i + block] = gram_block(X[i:i + block], Y, s, p, nconst)
    return out


def kmatw(X, Y, W, s, p, nconst, block=4096):
    """inner with Y.reduce = [LinearReduce(W.T)]: materialise the Gram block, then GEMM."""
    out = torch.empty((X.shape[0], W.shape[1]), dtype=torch.float32)
    for i in range(0, X.shape[0], block):
        g = gram_block(X[i:i + block], Y, s, p, nconst).to(torch.float32)   # .astype(float), vector.py:72
        out[i:i + block] = (W.T @ g.T).T                                    # Reduce.apply(axis=1), lincomb.py:140
    return out
