"""N > 1 path on CPU: world_size 2 over gloo.  Each rank owns a row block of X; results must equal
the single-process lowering for every foldable X-side chain, with and without an exchange."""
import os
import socket
import subprocess
import sys

import numpy as np

from jaxrk_b200 import sharded

HERE = os.path.dirname(os.path.abspath(__file__))


def test_row_block_partition():
    for n, world, align in ((24, 2, 1), (24, 2, 4), (25, 4, 1), (4096 * 3, 8, 4096), (7, 8, 1)):
        blocks = [sharded.row_block(n, r, world, align) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a, b), (c, d) in zip(blocks, blocks[1:]):
            assert b == c and a <= b
        assert all(a % align == 0 and b % align == 0 for a, b in blocks)
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= align


def test_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "dist_worker.py"), str(r), "2", str(port), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    for r in range(2):
        res = np.load(os.path.join(tmp_path, f"rank{r}.npy"), allow_pickle=True).item()
        assert len(res) == 9 * 3 + 3
        for name, err in res.items():
            assert err < 2e-6, (r, name, err)
