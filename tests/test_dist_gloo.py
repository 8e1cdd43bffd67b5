"""CPU, world_size 2 over gloo: the partitioning / gather / image-reduce logic of the
multi-GPU paths with CPU stand-ins for the CUDA compute callbacks."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from synthobs_b200 import dist as sdist


def test_lpt_sharding_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    sizes = (10 ** rng.uniform(3, 5, 1000)).astype(np.int64)
    for ws in (1, 2, 4, 8):
        shards = sdist.shard_galaxies(sizes, ws)
        allg = sorted(g for s in shards for g in s)
        assert allg == list(range(1000))
        loads = np.array([sizes[s].sum() for s in shards])
        assert loads.max() <= loads.mean() * 1.02
    assert sdist.shard_galaxies([5, 5, 5], 2) == sdist.shard_galaxies([5, 5, 5], 2)    # deterministic


def test_subset_csr_and_blocks():
    off = np.array([0, 3, 3, 7, 12])
    arr = {'a': np.arange(12.0)}
    sub, no = sdist.subset_csr(arr, off, [0, 2])
    assert np.array_equal(sub['a'], [0, 1, 2, 3, 4, 5, 6]) and np.array_equal(no, [0, 3, 7])
    sub, no = sdist.subset_csr(arr, off, [])
    assert sub['a'].size == 0 and np.array_equal(no, [0])
    cover = []
    for r in range(3):
        b, e = sdist.particle_block(10, r, 3)
        cover += list(range(b, e))
    assert cover == list(range(10))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws),
                      LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        rng = np.random.default_rng(5)
        sizes = rng.integers(0, 50, 37)
        offsets = np.concatenate([[0], np.cumsum(sizes)])
        vals = rng.random(int(offsets[-1]))

        def compute(indices):      # stand-in for the CUDA broadband kernel: per-galaxy [sum, count]
            return torch.tensor([[vals[offsets[g]:offsets[g + 1]].sum(), float(sizes[g])] for g in indices],
                                dtype=torch.float64).reshape(-1, 2)

        full = sdist.sharded_galaxy_map(compute, offsets)
        ref = np.array([[vals[offsets[g]:offsets[g + 1]].sum(), sizes[g]] for g in range(37)])
        ok1 = np.allclose(full.numpy(), ref, rtol=0, atol=0)
        # particle-block image reduce: each rank deposits its block of a 1-D "image"
        n = 1001
        pix = rng.integers(0, 16, n)
        w = rng.random(n)
        b, e = sdist.particle_block(n)
        img = torch.zeros(16, dtype=torch.float64)
        img.index_add_(0, torch.from_numpy(pix[b:e]), torch.from_numpy(w[b:e]))
        sdist.reduce_image(img)
        ok2 = np.allclose(img.numpy(), np.bincount(pix, weights=w, minlength=16), rtol=1e-13)

        # rows with more than one trailing axis (spectra [g, 5, n_lam], images [g, ndim, ndim]), float32
        def compute3(indices):
            return torch.tensor([[[g, sizes[g]], [2 * g, 1], [0, -g]] for g in indices], dtype=torch.float32).reshape(-1, 3, 2)

        full3 = sdist.sharded_galaxy_map(compute3, offsets)
        ref3 = np.array([[[g, sizes[g]], [2 * g, 1], [0, -g]] for g in range(37)], dtype=np.float32)
        ok1 = ok1 and full3.shape == (37, 3, 2) and np.array_equal(full3.numpy(), ref3)
        # fewer galaxies than ranks: the rank without a share returns an empty block and still takes part in the gather
        one_off = np.array([0, 5])
        mine = sdist.shard_galaxies(np.diff(one_off), ws)[rank]
        full1 = sdist.sharded_galaxy_map(lambda idx: torch.full((len(idx), 3), 7.0, dtype=torch.float64), one_off)
        ok1 = ok1 and full1.shape == (1, 3) and bool((full1 == 7.0).all()) and len(mine) == (1 if rank == 0 else 0)
        # one giant galaxy in particle blocks: per-rank partial sums + all-reduce == the whole
        blk = torch.tensor([vals[slice(*sdist.particle_block(vals.size))].sum()], dtype=torch.float64)
        dist.all_reduce(blk)
        ok2 = ok2 and np.isclose(float(blk), vals.sum(), rtol=1e-13)
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res
