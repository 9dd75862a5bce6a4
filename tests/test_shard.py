"""The N > 1 layout on CPU: world_size-2 gloo process group, each rank owning a
block of objects, result records gathered in global order (no GPU needed)."""
import os
import subprocess
import sys

import numpy as np

from rvspecfit_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from rvspecfit_b200 import shard
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
nobj = 11
a, b = shard.block_range(nobj, rank, world)
local = np.stack([np.arange(a, b) * 1.0, np.arange(a, b) ** 2 + 0.5, np.full(b - a, rank)], axis=1)
full = shard.gather_records(local, nobj)
assert full.shape == (nobj, 3)
assert np.array_equal(full[:, 0], np.arange(nobj)) and np.array_equal(full[:, 1], np.arange(nobj) ** 2 + 0.5)
assert list(full[:, 2]) == [0] * 6 + [1] * 5
dist.barrier()
dist.destroy_process_group()
print('rank', rank, 'ok')
'''


def test_block_ranges_cover_everything():
    for nobj in (0, 1, 7, 8, 1000, 100003):
        for world in (1, 2, 3, 8):
            blocks = [shard.block_range(nobj, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == nobj
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    assert np.array_equal(shard.gather_records(np.ones((3, 2)), 3), np.ones((3, 2)))


def test_two_rank_gloo_gather(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', MASTER_ADDR='127.0.0.1',
                   MASTER_PORT='29531', LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for rank, p in enumerate(procs):
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out
        assert f'rank {rank} ok' in out
