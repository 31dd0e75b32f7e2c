"""world_size-2 gloo tests (CPU) of the sweep partition + gather logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from femwell_b200.sweep import block_partition, group_index, pack_complex, sweep, unpack_complex


def test_block_partition_covers_everything():
    for n in (0, 1, 7, 256, 513):
        for world in (1, 2, 3, 8):
            blocks = [block_partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip_and_group_index():
    z = np.array([1 + 2j, 3 - 4j])
    assert np.array_equal(unpack_complex(np.array(pack_complex(z)), 2), z)
    wl = np.linspace(1.2, 1.9, 50)
    n = 2.0 - 0.3 * wl
    assert np.allclose(group_index(wl, n), 2.0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    points = [1.2 + 0.01 * i for i in range(11)]
    seen = []

    def solve_point(i, wl):
        seen.append(i)
        if i == 4:
            raise RuntimeError("synthetic failure")
        return pack_complex([wl * (1 + 1j), 2 * wl])

    table = sweep(points, solve_point, 4)
    np.save(os.path.join(out_dir, f"t{rank}.npy"), table)
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.array(seen))
    dist.destroy_process_group()


def test_sweep_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    t0, t1 = np.load(tmp_path / "t0.npy"), np.load(tmp_path / "t1.npy")
    assert np.array_equal(np.isnan(t0), np.isnan(t1)) and np.allclose(np.nan_to_num(t0), np.nan_to_num(t1))
    s0, s1 = np.load(tmp_path / "s0.npy"), np.load(tmp_path / "s1.npy")
    assert sorted(s0.tolist() + s1.tolist()) == list(range(11)) and not set(s0) & set(s1)
    assert t0.shape == (11, 5)
    assert t0[4, 4] == 1.0 and np.isnan(t0[4, :4]).all()
    ok = np.delete(np.arange(11), 4)
    wl = 1.2 + 0.01 * ok
    z = np.stack([unpack_complex(r, 2) for r in t0[ok]])
    assert np.allclose(z[:, 0], wl * (1 + 1j)) and np.allclose(z[:, 1], 2 * wl)
    assert (t0[ok, 4] == 0).all()


def test_sweep_without_process_group():
    table = sweep([1, 2, 3], lambda i, p: [p * 2.0], 1)
    assert table.shape == (3, 2) and np.allclose(table[:, 0], [2, 4, 6])
