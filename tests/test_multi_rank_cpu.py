"""N > 1 host logic on CPU: two `gloo` ranks (world_size 2, 127.0.0.1) each ask the C ABI for their
share of a frame's 32x32 super-tiles; together the shares must cover every tile exactly once, for
full frames, ragged sizes and column-strip windows; and the whole-frame dealing of a batch must give every
frame slot to exactly one rank.  No GPU work happens here."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch
import torch.distributed as dist
from raytracing_rb_b200 import deal_frames, tile_partition

dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
rank, world = dist.get_rank(), dist.get_world_size()
cases = [(1920, 1080, None), (3840, 2160, None), (192, 108, None), (33, 65, None), (1, 1, None),
         (1920, 1080, (480, 0, 960, 1080)), (200, 120, (50, 10, 51, 11))]
for (w, h, win) in cases:
    mine = tile_partition(w, h, rank, world, win)
    assert mine == sorted(mine) and len(set(mine)) == len(mine)
    stx = (w + 31) // 32
    t = torch.zeros(((h + 31) // 32) * stx, dtype=torch.int32)
    t[torch.tensor(mine, dtype=torch.long)] += 1 if mine else 0
    dist.all_reduce(t)
    x0, y0, x1, y1 = win or (0, 0, w, h)
    want = torch.zeros_like(t)
    for ty in range(y0 // 32, (y1 - 1) // 32 + 1):
        for tx in range(x0 // 32, (x1 - 1) // 32 + 1):
            want[ty * stx + tx] = 1
    assert torch.equal(t, want), (w, h, win)
    # balance: round-robin dealing differs by at most one tile between ranks
    n = torch.tensor([len(mine)]); lo = n.clone(); hi = n.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert int(hi) - int(lo) <= 1
# whole-frame dealing of a batch (bench.py's default at N > 1): every frame of the step is rendered by exactly one
# rank and lands in its own slot of rank 0's framebuffer
for B in (1, 16, 5):
    mine = deal_frames(rank, world, B)
    t = torch.zeros(world * B, dtype=torch.int32)
    for frame, slot in mine:
        assert frame == slot
        t[slot] += 1
    dist.all_reduce(t)
    assert torch.equal(t, torch.ones_like(t)), B
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
''' % ROOT


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_gloo_ranks_partition_tiles(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "raytracing_rb_b200", "csrc"), "-j4", "-s", "all"])
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % rank in o
