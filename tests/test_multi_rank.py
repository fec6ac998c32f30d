"""CPU-only, world_size = 2 over gloo: the replication sharding + single all-gather of
dynamic_factor_models_b200/replicate.py give the same records as a single rank.  The compute runs on
the host-emulation harness (tests/emu) because there is no GPU here; the GPU path uses the very same
driver with the CUDA library and NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp, what):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from dynamic_factor_models_b200 import Library, replicate
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lib = Library(build_emu.build())
    if what == "mc":
        rec = replicate.monte_carlo_em(lib, 5, 16, 2, 40, em_iters=3, rank=rank, world=world)
    else:
        m = _fitted_model(lib)
        rec, _ = replicate.bootstrap_irf(lib, m, 5, H=6, rank=rank, world=world)
    np.save(os.path.join(tmp, f"{what}_rank{rank}.npy"), rec)
    dist.destroy_process_group()
    lib.close()


def _fitted_model(lib):
    import dynamic_factor_models_b200 as D
    from dynamic_factor_models_b200 import replicate
    X = replicate.simulate_panel(14, 2, 70, rep=7, lib=lib)
    X[3:9, 2] = np.nan
    m = D.DFMModel(X, np.r_[np.ones(12, int), np.zeros(2, int)], 20, 20, 1, 70, 0, 2, 1e-8, 2, 2)
    D.estimate(m, lib=lib)
    return m


@pytest.mark.parametrize("what", ["mc", "boot"])
def test_two_ranks_match_one(tmp_path, what):
    import build_emu
    from dynamic_factor_models_b200 import Library, replicate
    lib = Library(build_emu.build())
    if what == "mc":
        ref = replicate.monte_carlo_em(lib, 5, 16, 2, 40, em_iters=3)
    else:
        ref, bands = replicate.bootstrap_irf(lib, _fitted_model(lib), 5, H=6)
        assert bands[50].shape == (2, 6, 2) and np.isfinite(bands[50]).all()
    lib.close()
    port = 29500 + os.getpid() % 2000 + (0 if what == "mc" else 1)
    mp.spawn(_worker, args=(2, port, str(tmp_path), what), nprocs=2, join=True)
    for rank in range(2):
        got = np.load(tmp_path / f"{what}_rank{rank}.npy")
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)


def test_shards_cover_and_are_gpu_count_independent():
    """Replication b's panel depends on b only: any batch split / shard gives the same bits, and the device stream is
    the one oracle/dgp.py restates."""
    import build_emu
    from dynamic_factor_models_b200 import Library, replicate
    lib = Library(build_emu.build())
    whole = lib.simulate_panels(0, 6, 8, 2, 20, replicate.SEED)
    for world in (2, 3):
        parts = []
        for rank in range(world):
            b, e = lib.shard_range(6, rank, world)
            parts.append(lib.simulate_panels(b, e - b, 8, 2, 20, replicate.SEED))
        assert np.array_equal(np.concatenate(parts), whole)
    from oracle.dgp import simulate_panel_device_stream
    np.testing.assert_allclose(replicate.simulate_panel(8, 2, 20, rep=3, lib=lib), simulate_panel_device_stream(8, 2, 20, rep=3)[0],
                               rtol=1e-10, atol=1e-12)
    lib.close()
