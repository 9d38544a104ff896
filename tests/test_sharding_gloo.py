"""CPU, world_size 2 and 3 over gloo: the N>1 path's host-side logic (shard ranges, allgatherv,
rank order == Walk order, unique-id broadcast)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,n", [(2, 50_001), (3, 7), (2, 0)])
def test_sharded_scan_over_gloo(world, n):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    port = 29500 + (os.getpid() + world * 7 + n) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(HERE, "_gloo_worker.py"), str(n)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "gloo-ok world=%d n=%d" % (world, n) in r.stdout
