"""Two-rank NCCL sharding on real GPUs (skipped when fewer than two devices are visible)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as td

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from emulator_b200 import dist, synthetic
    from emulator_b200.emulators import GaussianProcessEmulator, GaussianProcessEmulatorBins
    from emulator_b200.sensitivity import CrossCheck, sweep_predict

    spec, params, values = synthetic.make_model(3, 48, 6, seed=41)
    bins = GaussianProcessEmulatorBins(device=rank)
    bins.fit_model(spec, params, values)
    np.save(os.path.join(out_dir, f"bins_{rank}.npy"), bins.bin_fit_parameters)

    spec2, params2, values2 = synthetic.make_model(2, 8, 6, seed=42)
    cc = CrossCheck(device=rank)
    cc.build_emulators(spec2, params2, values2)
    total, _ = cc.get_mean_squared()
    np.save(os.path.join(out_dir, f"cc_{rank}.npy"), np.append(cc.cross_fit_parameters.ravel(), total))

    emu = GaussianProcessEmulator(device=rank)
    emu.fit_model(spec2, params2, values2)
    thetas = np.random.default_rng(0).random((37, 2))
    mean, var = sweep_predict(emu, thetas, np.linspace(0, 1, 5))
    np.save(os.path.join(out_dir, f"sweep_{rank}.npy"), np.stack([mean, var]))
    # the device default: no device argument anywhere -> this rank's GPU
    auto = GaussianProcessEmulator()
    auto.fit_model(spec2, params2, values2)
    assert auto.emulator.device == rank
    # walkers sharded over the ranks: every rank runs the same seeded sampler and evaluates its slice
    from emulator_b200.emulators import GaussianProcessEmulatorMCMC

    mcmc = GaussianProcessEmulatorMCMC(burn_in_steps=3, mcmc_steps=5, walkers=8, seed=4)
    mcmc.fit_model(spec2, params2, values2)
    np.save(os.path.join(out_dir, f"chain_{rank}.npy"), mcmc.hyperparameter_samples)
    # optimiser restarts shared out over the ranks: the same winner everywhere
    multi = GaussianProcessEmulator(restarts=3, restart_seed=1)
    multi.fit_model(spec2, params2, values2)
    np.save(os.path.join(out_dir, f"restarts_{rank}.npy"), np.append(multi.fit_result.restart_funs, multi.emulator.get_parameter_vector()))
    td.barrier()
    td.destroy_process_group()


def test_two_gpu_sharding_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    b0, b1 = np.load(tmp_path / "bins_0.npy"), np.load(tmp_path / "bins_1.npy")
    assert np.array_equal(b0, b1) and np.all(np.any(b0 != 0, axis=1))
    c0, c1 = np.load(tmp_path / "cc_0.npy"), np.load(tmp_path / "cc_1.npy")
    assert np.array_equal(c0[:-1], c1[:-1]) and c0[-1] == pytest.approx(c1[-1], rel=1e-9)
    s0, s1 = np.load(tmp_path / "sweep_0.npy"), np.load(tmp_path / "sweep_1.npy")
    assert s0.shape == (2, 37, 5) and np.array_equal(s0, s1)
    # the sharded chain is the single-process chain (same seed, deterministic kernels), on both ranks
    from emulator_b200 import synthetic
    from emulator_b200.emulators import GaussianProcessEmulatorMCMC

    spec2, params2, values2 = synthetic.make_model(2, 8, 6, seed=42)
    alone = GaussianProcessEmulatorMCMC(burn_in_steps=3, mcmc_steps=5, walkers=8, seed=4, device=0)
    alone.fit_model(spec2, params2, values2)
    k0, k1 = np.load(tmp_path / "chain_0.npy"), np.load(tmp_path / "chain_1.npy")
    assert np.array_equal(k0, k1) and np.array_equal(k0, alone.hyperparameter_samples)
    r0, r1 = np.load(tmp_path / "restarts_0.npy"), np.load(tmp_path / "restarts_1.npy")
    assert np.array_equal(r0, r1)
