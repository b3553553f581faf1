"""N>1 path on CPU: two gloo ranks shard a problem list (no data-path collective), 'solve' their
shards with the oracle, and only max-reduce the time / sum the rows, as bench.py does on GPUs."""
import os
import socket
import sys
import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from peaksegdisk_b200 import shard, synth
    import oracle_bind
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [300 + 37 * (k % 11) for k in range(23)]
    mine = shard.lpt_assign(sizes, world)[rank]
    rows, peaks = 0, []
    for i in mine:
        s, e, c = synth.poisson_problem(1000 + i, sizes[i])
        st, summ, _ = oracle_bind.solve_rows(s, e, c, 10.0)
        assert st == 0
        rows += len(c); peaks.append((i, int(summ[2])))
    ms, total = shard.reduce_time_and_rows(10.0 + rank, rows, dist)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([ms, total, rows] + [v for p in peaks for v in p]))
    assert shard.rank_seeds(rank, 4) == [rank * 1024 + k for k in range(4)]
    dist.destroy_process_group()


def test_two_ranks_shard_without_data_collectives(tmp_path):
    from peaksegdisk_b200 import shard, synth
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0[0] == r1[0] == 11.0                      # MAX over ranks
    assert r0[1] == r1[1] == r0[2] + r1[2]             # SUM of rows
    ids = sorted(int(v) for v in np.concatenate((r0[3::2], r1[3::2])))
    assert ids == list(range(23))                      # every problem solved exactly once
    sizes = [300 + 37 * (k % 11) for k in range(23)]
    loads = [sum(sizes[i] for i in sh) for sh in shard.lpt_assign(sizes, world)]
    assert abs(loads[0] - loads[1]) <= max(sizes)      # balanced


def test_lpt_assign_properties():
    from peaksegdisk_b200 import shard
    rng = np.random.default_rng(0)
    sizes = rng.integers(10, 100000, size=101).tolist()
    for world in (1, 2, 4, 8):
        sh = shard.lpt_assign(sizes, world)
        assert sorted(i for s in sh for i in s) == list(range(101))
        loads = [sum(sizes[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(sizes)
