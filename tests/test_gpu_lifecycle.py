"""Handle life cycle on the GPU: repeated create/destroy leaks no device memory, several handles coexist, an error does not
poison later calls, L0 calls work on the caller's (non-default) stream."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free():
    import torch
    torch.cuda.synchronize()
    return torch.cuda.mem_get_info()[0]


def test_create_destroy_does_not_leak(b2s, gpu):
    from b200stencil import part1, part2
    # first use allocates library scratch / CUDA module memory once
    g = part1.Diffusion3D(64, 64, 64); g.init_gaussian(); g.iterate(2); g.close()
    m = part2.MGHandle(257, 257); m.close()
    def one_round():
        g = part1.Diffusion3D(96, 64, 48, nslabs=2, devices=[0, 0]); g.init_gaussian(); g.iterate(3); g.close()
        m = part2.MGHandle(513, 257)
        x = part2.zeros(513, 257)
        m.solve(x, part2.to_device(np.ones((513, 257))), 1.0 / 256, 0.0, 1e-6, 3, False)
        m.close()
        del x
        s = part2.NavierStokes2D(part2.SimIn_t(nx=129, ny=33, beta=0.5, niters=3)); s.init_cosine("T"); s.step(); s.close()
    one_round()  # torch's caching allocator and CUDA's graph/module pools reach their steady state
    import torch
    torch.cuda.empty_cache()
    base = _free()
    for _ in range(10):
        one_round()
    torch.cuda.empty_cache()
    assert base - _free() < 64 * 2 ** 20, (base, _free())  # nothing (beyond allocator noise) stays allocated


def test_handles_coexist_and_errors_do_not_poison(b2s, gpu, oracle):
    from b200stencil import capi, part1, part2
    a = part1.Diffusion3D(64, 64, 64)
    b = part1.Diffusion3D(32, 32, 32)
    a.init_gaussian(); b.init_gaussian()
    with pytest.raises(capi.B2SError):
        part1.Diffusion3D(64, 64, 64, nslabs=2, devices=[0, 99])  # no such device
    with pytest.raises(capi.B2SError):
        part2.MGHandle(100, 100)
    ea, eb = a.iterate(7), b.iterate(7)
    oa, ob = oracle.Diffusion3D(64, 64, 64), oracle.Diffusion3D(32, 32, 32)
    assert np.allclose(ea, oa.iterate(7), rtol=1e-12, atol=0) and np.allclose(eb, ob.iterate(7), rtol=1e-12, atol=0)
    assert np.array_equal(a.get("Htau"), oa.get("Htau")) and np.array_equal(b.get("Htau"), ob.get("Htau"))
    a.close(); b.close()


def test_l0_calls_on_a_side_stream(b2s, gpu, oracle):
    import torch
    from b200stencil import part2
    shape = (129, 65)
    u = np.asfortranarray(np.random.default_rng(3).random(shape)); f = np.asfortranarray(np.random.default_rng(4).random(shape))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        du, df, dres = part2.to_device(u), part2.to_device(f), part2.zeros(*shape)
        r = part2.iteration_2DPoisson(du, df, 1.0 / 64, 1.5, dres)
    st.synchronize()
    uo, reso = u.copy(order="F"), oracle.farray(shape)
    r_o = oracle.jacobi2d(uo, f, 1.0 / 64, 1.5, reso)
    assert np.array_equal(part2.to_host(du), uo) and abs(r - r_o) <= 1e-12 * r_o


def test_reference_experiment_glue_writes_reference_csvs(b2s, gpu, tmp_path):
    """f4: part1_scaling_experiments.jl / multigrid_bench.jl rows in the reference's CSV schemas (tiny shapes here); the
    Work column still decodes to the timed PT iteration count like the published files do."""
    import pandas as pd
    from b200stencil import experiments as E
    fn = str(tmp_path / "bench_diffusion_scaling_gpu.csv")
    rows = E.part1_scaling_experiments(n_mpi_ranks=2, filename=fn, devices=[0, 0], n_global=32, ttot=1.0, tol=1e-4,
                                       layouts=("reference", "zslab"))
    df = pd.read_csv(fn)
    assert list(df.columns) == E.SCALING_COLUMNS and len(df) == 4
    for r in rows:
        cells = (r["local_grid"][0] - 2) * (r["local_grid"][1] - 2) * (r["local_grid"][2] - 2)
        assert r["Work"] == 2.0 * r["timed_iters"] * 27 * cells
    # x-split (reference layout) and z-split give the same counts by the symmetry of the problem (lag-2 halos)
    by = {(r["layout"], r["strong_scaling"], r["use_shared_memory"]): r["timed_iters"] for r in rows}
    for ss in (True, False):
        assert by[("reference", ss, True)] == by[("reference", ss, False)] == by[("zslab", ss, True)]
    fn2 = str(tmp_path / "bench_multigrid_gpu.csv")
    E.multigrid_bench(ks=(7,), filename=fn2, samples=2)
    df2 = pd.read_csv(fn2)
    assert list(df2.columns) == E.MULTIGRID_COLUMNS and len(df2) == 2 * 2 * 2 and (df2.median_time > 0).all()
    assert set(df2.coarse_solver) == {"jacobi", "conjugate_gradient"} and set(df2.l) == {2, 3}
