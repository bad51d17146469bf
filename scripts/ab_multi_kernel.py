"""A/B on ONE GPU: (a) the single-slab step kernel, lean vs exchange-capable instantiation (B2S_FORCE_MULTI_KERNEL=1 in the
environment selects the latter); (b) two z-slabs of 512x512x258 hosted on the same GPU (flags, pushes through local memory,
lagged norm: everything of the multi-GPU protocol except NVLink) against one 512^3 slab. Prints ms per iteration per slab."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200stencil  # noqa
from b200stencil import part1
out = {"force_multi": os.environ.get("B2S_FORCE_MULTI_KERNEL", "0")}
for name, kw, cells in (("one_slab_512", dict(nz=512), 510.0 ** 3),
                        ("two_slabs_258_same_gpu", dict(nz=258, nslabs=2, devices=[0, 0], scale_physical_size=True), 2 * 510.0 * 510 * 256)):
    g = part1.Diffusion3D(512, 512, **kw)
    g.init_gaussian()
    for _ in range(3):
        g.iterate(100, want_hist=False)
    ms = 0.0
    for _ in range(5):
        g.iterate(200, want_hist=False)
        ms += g.stats()[1]
    out[name] = {"ms_per_iteration": ms / 1000, "T_eff_GBs": 24.0 * cells * 1000 / (ms * 1e-3) / 1e9}
    g.close()
print(json.dumps(out))
