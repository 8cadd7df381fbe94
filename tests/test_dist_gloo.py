"""Round-robin sharding + all-gather with two gloo ranks on CPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    from emulator_b200 import dist

    assert dist.rank() == rank and dist.world_size() == world
    count = 7
    mine = dist.my_indices(count)
    assert mine == list(range(rank, count, world))
    rows = np.zeros((count, 3))
    for i in mine:
        rows[i] = [i, 10 * i, rank]
    full = dist.all_gather_rows(rows, mine, count)
    expected = np.array([[i, 10 * i, i % world] for i in range(count)], dtype=float)
    assert np.array_equal(full, expected)

    start, stop = dist.contiguous_slice(11)
    block = np.arange(start, stop, dtype=float).reshape(-1, 1) * np.ones((1, 2))
    cat = dist.all_gather_concat(block, 11)
    assert np.array_equal(cat[:, 0], np.arange(11.0))

    # sharded bins emulator through the oracle-backed GP: both ranks end with all optima
    import oracle_george
    import emulator_b200.george_compat as gc

    gc.GP = oracle_george.OracleBackedGP
    from emulator_b200 import synthetic
    from emulator_b200.emulators import GaussianProcessEmulatorBins
    from emulator_b200.sensitivity import CrossCheck

    spec, params, values = synthetic.make_model(2, 16, 4, seed=31)
    emu = GaussianProcessEmulatorBins()
    emu.fit_model(spec, params, values)
    np.save(os.path.join(out_dir, f"bins_{rank}.npy"), emu.bin_fit_parameters)
    mu, _ = emu.predict_values(np.array(emu.bin_centers), {"p0": 0.4, "p1": 0.6})
    np.save(os.path.join(out_dir, f"binsmu_{rank}.npy"), mu)

    spec, params, values = synthetic.make_model(2, 6, 5, seed=32)
    cc = CrossCheck()
    cc.build_emulators(spec, params, values)
    total, _ = cc.get_mean_squared()
    np.save(os.path.join(out_dir, f"cc_{rank}.npy"), np.append(cc.cross_fit_parameters.ravel(), total))
    td.barrier()
    td.destroy_process_group()


def test_two_rank_sharding_and_gather(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    b0, b1 = np.load(tmp_path / "bins_0.npy"), np.load(tmp_path / "bins_1.npy")
    assert np.array_equal(b0, b1) and np.all(np.any(b0 != 0, axis=1))
    assert np.allclose(np.load(tmp_path / "binsmu_0.npy"), np.load(tmp_path / "binsmu_1.npy"), rtol=1e-12)
    c0, c1 = np.load(tmp_path / "cc_0.npy"), np.load(tmp_path / "cc_1.npy")
    assert np.array_equal(c0[:-1], c1[:-1]) and c0[-1] == pytest.approx(c1[-1], rel=1e-12)


def test_single_process_degenerates_to_one_rank():
    from emulator_b200 import dist

    assert dist.world_size() == 1 and dist.my_indices(5) == [0, 1, 2, 3, 4]
    rows = np.arange(6.0).reshape(3, 2)
    assert np.array_equal(dist.all_gather_rows(rows, [0, 1, 2], 3), rows)
    assert dist.contiguous_slice(10, 1, 3) == (4, 7)
