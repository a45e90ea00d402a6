"""CPU oracle of the PanFusion denoise hot path — TEST INFRASTRUCTURE ONLY.

A restatement (numpy for the float64 grid math, plain PyTorch fp32 for everything else) of the reference's
algorithm, each function citing the reference file:line it follows. Nothing under `panfusion_b200/` imports
this package: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may use it, and only as the checker or the timed CPU baseline — never as the product path.

Pinning status ("parity pinned by reference execution"): the reference repo has no tests and no golden
vectors. In the build container the reference's own files are executed by path with
third-party stand-ins (`oracle/ref_loader.py`) and compared with this restatement by
`oracle/make_golden.py`, which also writes the fixtures under `tests/golden/`. Third-party pieces whose
source is absent (kornia 0.7.2 remap / create_meshgrid / gaussian_blur2d, xformers 0.0.22
memory_efficient_attention, diffusers 0.24.0 UNet blocks + DDIM) are restated from their published
behaviour and are therefore "parity unpinned" beyond the reference's call sites.
"""
