"""Row-sharded path on real GPUs over NCCL (needs >= 2 GPUs: `gpurun --gpus 2`); skipped on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(sys.argv[0]))) if False else os.getcwd())
rank, world, port = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from jaxrk_b200 import sharded
from jaxrk_b200.kern import GenGaussKernel
from jaxrk_b200.reduce import BalancedRed, LinearReduce, Mean, Prefactors
from jaxrk_b200.rkhs import FiniteVec, inner
rng = np.random.RandomState(0)
X, Y = rng.randn(4096, 16).astype(np.float32), rng.randn(3000, 16).astype(np.float32)
W, pp = rng.randn(8, 3000).astype(np.float32), rng.rand(4096).astype(np.float32)
k = GenGaussKernel.make_gauss(4.0)
yv = FiniteVec(k, Y, [LinearReduce(W)])
for cx in ([], [Mean()], [Prefactors(pp), BalancedRed(512, True)], [LinearReduce(rng.randn(4, 4096).astype(np.float32))]):
    got = sharded.sharded_inner(FiniteVec(k, X, cx), yv)
    want = inner(FiniteVec(k, X, cx), yv)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=2e-5, atol=1e-6), (rank, float((got - want).abs().max()))
s, e = sharded.row_block(4096, rank, world)
G = sharded.sharded_gram(k, torch.from_numpy(X[s:e]).cuda(), torch.from_numpy(Y).cuda())
assert torch.equal(G, k(X, Y)[s:e])
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return sorted({2, min(n, 8)} - {0, 1}) if n >= 2 else [2]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", _worlds())
def test_nccl_world(tmp_path, world):
    """world = 2 and, on a bigger box, every GPU of the node (gpurun --gpus 8): sharded_inner's all_reduce / all_gather and
    sharded_gram over NCCL against the single-GPU result."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = os.path.dirname(HERE)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(port)], cwd=root, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
