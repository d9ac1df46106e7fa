"""CPU, world_size 2, gloo: the multi-GPU host logic -- contiguous sharding of
pairs over ranks and the gather of per-pair results."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_pairs, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from gstpeaq_b200 import parallel
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = parallel.shard_range(n_pairs, rank, world)
    # stand-in for the engine: result row p carries f(p) so placement can be checked
    local = np.zeros(count, dtype=parallel.RESULT_DTYPE)
    local["odg"] = -(np.arange(first, first + count) % 7) * 0.5
    local["movs"][:, 0] = np.arange(first, first + count)
    local["frames_fft"] = 468
    full = parallel.gather_results(local, n_pairs, device=torch.device("cpu"))
    q.put((rank, first, count, full["odg"].copy(), full["movs"][:, 0].copy(), full["frames_fft"].copy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [8, 7, 1])
def test_shard_and_gather_world_size_2(n_pairs):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + n_pairs
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, first, count, odg, m0, frames in got:
        covered += list(range(first, first + count))
        np.testing.assert_array_equal(m0, np.arange(n_pairs))
        np.testing.assert_array_equal(odg, -(np.arange(n_pairs) % 7) * 0.5)
        np.testing.assert_array_equal(frames, 468)
    assert sorted(covered) == list(range(n_pairs))


def test_shard_range_is_contiguous_and_balanced():
    sys.path.insert(0, ROOT)
    from gstpeaq_b200 import parallel
    for n in (0, 1, 7, 8, 4096, 65536, 65537):
        for w in (1, 2, 4, 8):
            pos = 0
            sizes = []
            for r in range(w):
                f, c = parallel.shard_range(n, r, w)
                assert f == pos
                pos += c
                sizes.append(c)
            assert pos == n and max(sizes) - min(sizes) <= 1
