"""Multi-GPU parity of hot path 1 (needs >= 2 GPUs; skipped on a single-GPU box): z-slabs on different devices driven
in-process, and one process per GPU with CUDA-IPC peer mappings (the layout bench.py uses under torchrun).
Two ranks are never run on ONE GPU: their kernels wait on each other (B200_PROFILING.md)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import b200stencil  # noqa: F401
    from b200stencil import capi
    return capi.device_count()


@pytest.mark.parametrize("halo_mode", [0, 1])
def test_inprocess_two_devices_match_rank_emulation(b2s, gpu, oracle, halo_mode):
    if gpu < 2:
        pytest.skip("needs 2 GPUs")
    from b200stencil import part1
    shape, N = (64, 64, 34), 2
    o = oracle.Diffusion3D(*shape, dims=(1, 1, N), halo_mode=halo_mode)
    g = part1.Diffusion3D(*shape, nslabs=N, devices=[0, 1], halo_mode=halo_mode)
    g.init_gaussian()
    for chunk in (1, 2, 3, 30):
        eo, eg = o.iterate(chunk), g.iterate(chunk)
        assert np.allclose(eg, eo, rtol=1e-12, atol=0)
        for r in range(N):
            assert np.array_equal(g.get("Htau", r), o.get("Htau", r))
    assert g.solve_timestep(1e-6)[0] == o.solve_timestep(1e-6)[0]
    g.close()


@pytest.mark.parametrize("halo_mode", [0, 1])
@pytest.mark.parametrize("shape,dims", [((64, 32, 18), (2, 2, 1)), ((20, 18, 16), (2, 2, 2))])
def test_inprocess_general_decomposition_on_two_devices(b2s, gpu, oracle, halo_mode, shape, dims):
    """A 2x2x1 / 2x2x2 rank grid spread over two GPUs (ranks alternate between the devices): the per-axis plane copies
    cross the NVLink and are ordered by events between the device streams."""
    if gpu < 2:
        pytest.skip("needs 2 GPUs")
    from b200stencil import part1
    nr = dims[0] * dims[1] * dims[2]
    o = oracle.Diffusion3D(*shape, dims=dims, halo_mode=halo_mode)
    g = part1.Diffusion3D(*shape, dims=dims, devices=[r % 2 for r in range(nr)], halo_mode=halo_mode)
    g.init_gaussian()
    for chunk in (1, 2, 3, 30):
        eo, eg = o.iterate(chunk), g.iterate(chunk)
        assert np.allclose(eg, eo, rtol=1e-12, atol=0)
        for r in range(nr):
            assert np.array_equal(g.get("Htau", r), o.get("Htau", r)), r
    assert g.solve_timestep(1e-6)[0] == o.solve_timestep(1e-6)[0]
    assert np.array_equal(g.gather(), o.gather())
    g.close()


@pytest.mark.parametrize("args", [["64", "64", "34", "0", "tma"], ["32", "32", "18", "1", "direct"],
                                  ["64", "20", "18", "0", "tma", "x"], ["20", "18", "16", "1", "direct", "y"],
                                  ["64", "32", "18", "0", "tma", "xy"]])
def test_one_process_per_gpu_matches_rank_emulation(b2s, gpu, args):
    """z-slabs, and general decompositions (split in x, in y, 2x2x1 when 4 GPUs are there) with one process per GPU: the
    plane copies pull from the neighbours' IPC-mapped arenas, phases are separated by the device-side rank barrier."""
    if gpu < 2:
        pytest.skip("needs 2 GPUs")
    n = min(gpu, 4)
    args = list(args)
    if len(args) == 6:
        kind = args.pop()
        if kind == "xy":
            if gpu < 4:
                pytest.skip("needs 4 GPUs")
            n, dims = 4, ["2", "2", "1"]
        else:
            n, dims = 2, (["2", "1", "1"] if kind == "x" else ["1", "2", "1"])
        args += dims
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tests", "mp_diffusion_check.py")] + args,
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for k in range(n):
        assert f"rank{k} ok" in r.stdout


@pytest.mark.parametrize("halo_mode", [0, 1])
@pytest.mark.parametrize("shape,dims", [((64, 20, 18), (2, 1, 1)), ((20, 34, 16), (1, 2, 1))])
def test_inprocess_general_decomposition_one_rank_per_device(b2s, gpu, oracle, halo_mode, shape, dims):
    """A general rank grid with every rank on its own GPU, driven from one process: update_halo! as pulls separated by the
    device-side rank barrier (no cross-device events) -- bit-exact against the oracle's rank emulation."""
    if gpu < 2:
        pytest.skip("needs 2 GPUs")
    from b200stencil import part1
    o = oracle.Diffusion3D(*shape, dims=dims, halo_mode=halo_mode)
    g = part1.Diffusion3D(*shape, dims=dims, devices=[0, 1], halo_mode=halo_mode)
    g.init_gaussian()
    for chunk in (1, 2, 3, 30):
        eo, eg = o.iterate(chunk), g.iterate(chunk)
        assert np.allclose(eg, eo, rtol=1e-12, atol=0)
        for r in range(2):
            assert np.array_equal(g.get("Htau", r), o.get("Htau", r)), r
    assert g.solve_timestep(1e-6)[0] == o.solve_timestep(1e-6)[0]
    assert np.array_equal(g.gather(), o.gather())
    g.close()
