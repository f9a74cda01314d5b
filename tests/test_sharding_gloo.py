"""world_size-2 (gloo, CPU) coverage of the multi-GPU host logic: band partition, the set of AC
groups a rank must receive (own rows + filter halo), padded all-gather assembly.  The per-rank compute
is done by the oracle here (tests may use it); on GPUs bench.py runs the same logic over NCCL."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import jxl_workload as wl
from libjxl_b200 import abi, sharding

pytestmark = pytest.mark.usefixtures("built")


def _worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cpu
    desc, coeffs = wl.synthetic_frame(300, 1000, seed=5)          # 2 x 4 groups
    bands = sharding.band_partition(desc.ysize_groups, world)
    y0g, nyg = bands[rank]
    need = sharding.groups_needed(desc, y0g, nyg)
    # the rank only ever sees the groups it needs
    mine = np.zeros_like(coeffs)
    mine[:, need] = coeffs[:, need]
    full = cpu.render_frame(desc, mine)
    y0, rows = sharding.band_pixel_rows(desc, y0g, nyg)
    max_rows = max(sharding.band_pixel_rows(desc, a, b)[1] for a, b in bands)
    slot = torch.zeros((max_rows, desc.xsize, 3), dtype=torch.float32)
    slot[:rows] = torch.from_numpy(full[y0:y0 + rows])
    gathered = torch.empty((world, max_rows, desc.xsize, 3), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered.view(-1), slot.view(-1))
    frame = np.concatenate([gathered[r, :sharding.band_pixel_rows(desc, *bands[r])[1]].numpy() for r in range(world)])
    if rank == 0:
        want = cpu.render_frame(desc, coeffs)
        q.put((bool(np.array_equal(frame, want)), frame.shape, len(need), desc.num_groups))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_band_sharding_assembles_exact_frame():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, shape, n_need, n_groups = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and shape == (1000, 300, 3)
    assert n_need < n_groups          # rank 0 did not need the far band's groups
