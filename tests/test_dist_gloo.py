"""N>1 host logic on CPU: batch sharding + the single all-gather of u* (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from pympc_b200.dist import shard_range


def test_shard_range_partitions_the_batch():
    for B in (1, 7, 64, 65536, 524288):
        for W in (1, 2, 3, 8):
            spans = [shard_range(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pympc_b200.dist import shard_range, allgather_outputs
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
B, nu = 10, 3
full = torch.zeros(B, nu, dtype=torch.float64)
s, e = shard_range(B, dist.get_rank(), 2)
full[s:e] = torch.arange(s, e, dtype=torch.float64)[:, None] + 0.25 * dist.get_rank()
out = allgather_outputs(full, s, e)
exp = torch.arange(B, dtype=torch.float64)[:, None].repeat(1, nu); exp[5:] += 0.25
assert torch.equal(out, exp), out
dist.destroy_process_group()
print("ok")
"""


def test_allgather_outputs_world_size_2(tmp_path):
    script = tmp_path / "w.py"; script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
