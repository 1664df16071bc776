"""CPU, world_size 2, gloo: the multi-rank plumbing of bench.py (one independent stream per rank, max-over-ranks
timing, rank-0-only reporting).  No GPU work is involved: streams are independent, there is no data-path collective."""
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    scene = bench.make_scene(bench.CONFIGS[3], rank)
    frames = [scene.frame(t) for t in range(2)]
    ms = bench.reduce_max_ms(10.0 + 5.0 * rank, world, torch.device("cpu"))
    dist.barrier()
    out[rank] = (float(scene.x0[0] + scene.vel[0, 0]), int(frames[1].sum() % 1000003), ms)
    dist.destroy_process_group()


def test_two_ranks_track_different_streams_and_agree_on_max_time():
    world = 2
    with mp.get_context("spawn").Manager() as m:      # no fork() of the multi-threaded pytest process
        out = m.dict()
        mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0[0] != r1[0] and r0[1] != r1[1], "each rank must get its own synthetic stream"
    assert r0[2] == r1[2] == 15.0, "timing is the max over ranks"


def test_reference_arm_prints_only_on_rank0():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "6"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
