"""world_size-2 CPU test (gloo) of the frame sharding + record all-gather used by the multi-GPU stream mode."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, F, cap, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("self_commit_orb-slam2_b200.sharding")
    total = world * F
    lo, hi = sh.shard_range(total, world, rank)
    assert hi - lo == F
    # deterministic fake records: frame f has count f+3 and bytes derived from f
    kps = torch.zeros((F, cap, 28), dtype=torch.uint8)
    desc = torch.zeros((F, cap, 32), dtype=torch.uint8)
    counts = torch.zeros(F, dtype=torch.int32)
    for i, f in enumerate(range(lo, hi)):
        kps[i] = (f * 7) % 251
        desc[i] = (f * 13) % 239
        counts[i] = f + 3
    g_kps = torch.zeros((world, F, cap, 28), dtype=torch.uint8)
    g_desc = torch.zeros((world, F, cap, 32), dtype=torch.uint8)
    g_counts = torch.zeros((world, F), dtype=torch.int32)
    sh.gather_records(kps, desc, counts, g_kps, g_desc, g_counts)
    ok = True
    for r in range(world):
        l2, _ = sh.shard_range(total, world, r)
        for i in range(F):
            f = l2 + i
            ok &= int(g_counts[r, i]) == f + 3
            ok &= int(g_kps[r, i, 0, 0]) == (f * 7) % 251 and int(g_desc[r, i, cap - 1, 31]) == (f * 13) % 239
    s_k = torch.zeros((cap, 28), dtype=torch.uint8)
    s_d = torch.zeros((cap, 32), dtype=torch.uint8)
    s_c = torch.zeros(1, dtype=torch.int32)
    sh.take_predecessor(g_kps, g_desc, g_counts, rank, world, s_k, s_d, s_c)
    pred = (lo - 1) % total  # ring-ordered stream
    ok &= int(s_c[0]) == pred + 3 and int(s_k[0, 0]) == (pred * 7) % 251 and int(s_d[0, 0]) == (pred * 13) % 239
    # the stereo depths travel the same way (gather_array): frame f carries depth f + 0.5 everywhere
    dp = torch.zeros((F, cap), dtype=torch.float32)
    for i, f in enumerate(range(lo, hi)):
        dp[i] = f + 0.5
    g_dp = torch.zeros((world, F, cap), dtype=torch.float32)
    sh.gather_array(dp, g_dp)
    ok &= float(g_dp[sh.predecessor_source(rank, world), F - 1, cap - 1]) == pred + 0.5
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gather_and_predecessor():
    world, F, cap = 2, 3, 16
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, F, cap, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


def test_shard_range_partition():
    sh = importlib.import_module("self_commit_orb-slam2_b200.sharding")
    for total in (7, 64, 4096):
        for world in (1, 2, 4, 8):
            got = []
            for r in range(world):
                lo, hi = sh.shard_range(total, world, r)
                got += list(range(lo, hi))
            assert got == list(range(total))
    assert sh.predecessor_source(0, 8) == 7 and sh.predecessor_source(3, 8) == 2
