"""Multi-GPU partitioning rules (oblas_b200/shard.py) incl. a world_size-2 gloo run."""
import os
import subprocess
import sys

from helpers import ROOT
from oblas_b200 import shard


def test_batch_ranges_cover_exactly():
    for n in (1, 7, 256, 1024, 4096, 4097):
        for w in (1, 2, 4, 8):
            rs = [shard.batch_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1


def test_column_slices_are_tile_aligned():
    for nbytes in (65536, 1280, 272, 10000 * 16):
        for w in (1, 2, 4, 8):
            sl = [shard.column_slice(nbytes, r, w) for r in range(w)]
            assert sl[0][0] == 0 and sl[-1][1] == nbytes
            for i in range(w - 1):
                assert sl[i][1] == sl[i + 1][0] and (sl[i][1] % 256 == 0 or sl[i][1] == nbytes)


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from oblas_b200 import shard
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"],
                        rank=int(os.environ["RANK"]), world_size=2)
r = dist.get_rank()
C = torch.arange(64, dtype=torch.uint8).reshape(8, 8) if r == 0 else torch.zeros(8, 8, dtype=torch.uint8)
shard.broadcast_coefficients(C, src=0)
assert C.sum().item() == sum(range(64))
a, b = shard.batch_range(9, r, 2)
local = torch.arange(a, b, dtype=torch.int32)
allr = shard.gather_ranks(local)
assert allr.tolist() == list(range(9)), allr
lo, hi = shard.column_slice(1280, r, 2)
t = torch.tensor([hi - lo]); dist.all_reduce(t); assert t.item() == 1280
dist.destroy_process_group()
print("ok", r)
""" % ROOT


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=240)
        assert p.returncode == 0, err
        assert "ok" in out
