"""Worker of the world_size-2 gloo tests (tests/test_sharded_gloo.py): runs on CPU with the oracle-backed stand-ins
of tests/test_lowering_cpu.py in place of the native calls."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import jaxrk_b200._array as arr
    from jaxrk_b200 import _native as nat
    from jaxrk_b200 import sharded
    from jaxrk_b200.kern import GenGaussKernel
    from jaxrk_b200.reduce import BalancedRed, Center, LinearReduce, Mean, Prefactors, Sum
    from jaxrk_b200.rkhs import FiniteVec, inner
    import test_lowering_cpu as stand

    arr._device = torch.device("cpu")
    nat.gram, nat.dist, nat.kmatw, nat.gram_groupsum = stand.fake_gram, stand.fake_dist, stand.fake_kmatw, stand.fake_groupsum
    nat.gram_affine = stand.fake_gram_affine

    rng = np.random.RandomState(7)
    P, Q = rng.randn(24, 4).astype(np.float32), rng.randn(36, 4).astype(np.float32)
    pp, A_p, A_q = rng.rand(24).astype(np.float32), rng.randn(3, 24).astype(np.float32), rng.randn(5, 36).astype(np.float32)
    k = GenGaussKernel.make(1.7, 1.5)
    chains = {"none": [], "pref": [Prefactors(pp)], "sum": [Sum()], "mean": [Mean()], "pref_mean": [Prefactors(pp), Mean()],
              "bal4": [BalancedRed(4)], "bal3avg_center": [BalancedRed(3, True), Center()], "lin": [LinearReduce(A_p)],
              "pref_lin_sum": [Prefactors(pp), LinearReduce(A_p), Sum()]}
    ychains = {"none": [], "lin": [LinearReduce(A_q)], "mean": [Mean()]}
    res = {}
    for nx, cx in chains.items():
        for ny, cy in ychains.items():
            got = sharded.sharded_inner(FiniteVec(k, P, cx), FiniteVec(k, Q, cy))
            want = inner(FiniteVec(k, P, cx), FiniteVec(k, Q, cy))   # single-process lowering on the same stand-ins
            res[f"{nx}__{ny}"] = float((got - want).abs().max() / (want.abs().max() + 1e-30))
            assert got.shape == want.shape, (nx, ny, got.shape, want.shape)
    # row-sharded Gram without any collective, and its gathered form
    s, e = sharded.row_block(24, rank, world)
    G_loc = sharded.sharded_gram(k, torch.from_numpy(P[s:e]), torch.from_numpy(Q))
    G_all = sharded.sharded_gram(k, torch.from_numpy(P[s:e]), torch.from_numpy(Q), gather=True)
    G = k(P, Q)
    res["gram_local"] = float((G_loc - G[s:e]).abs().max())
    res["gram_gather"] = float((G_all - G).abs().max())
    # unreduced K@W stays sharded: no exchange
    T_loc = sharded.sharded_inner_local(k, torch.from_numpy(P[s:e]), [], 24, s, FiniteVec(k, Q, [LinearReduce(A_q)]), gather=False)
    res["kmatw_local"] = float((T_loc - inner(FiniteVec(k, P), FiniteVec(k, Q, [LinearReduce(A_q)]))[s:e]).abs().max())
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), res, allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
