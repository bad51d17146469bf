"""b200-stencil: B200-native (sm_100a) kernels behind the entry points of ntselepidis/FinalProjectRepo.jl.

The directory name contains a dot, so it cannot be imported with a plain `import`; use the loader at the repo root:

    import b200stencil            # registers this package as the module `b200stencil`
    from b200stencil import part1, part2, capi

Only what the two hot paths need lives here: csrc/ (CUDA kernels + C ABI), _capi.py (ctypes binding of
include/b200stencil.h), part1.py / part2.py (host-side mirrors of the reference's scripts-part1 / scripts-part2 entry
points), dist.py (one-process-per-GPU plumbing over torch.distributed) and julia/ (the ccall shims).
"""
from . import _capi as capi  # noqa: F401

__all__ = ["capi"]
