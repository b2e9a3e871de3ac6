"""CPU oracle for the surrogate time-stepping hot path (TEST INFRASTRUCTURE ONLY).

This package restates, on the CPU, the arithmetic of the reference's rollout path
(`pytorch_networks_convae.py` TS / NewFluidNet / FluidLayer / ADNet and
`symmetric_layers_torch.py` SymmetricConv2d).  It exists so that the CUDA path can be
checked on machines where `/root/reference` is not present (the GPU boxes).

Rules (enforced by tests/test_layout.py):
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
    reference` leg may import anything from here -- and only as the checker / the timed CPU
    baseline, never as the product path;
  * nothing under `pbml_mantle_convection_b200/` imports `oracle`.

Parity pin: the reference ships no tests or golden vectors for this path,
so the oracle is pinned against outputs of the reference itself, generated in the build
container by `tests/golden/make_golden.py` (which imports the real reference modules) and
committed under `tests/golden/*.npz`.  `tests/test_oracle_golden.py` re-checks both oracle
restatements against those vectors on every run.

Two restatements:
  * `ref_numpy` -- index-level numpy (no torch ops); pins exact semantics of every operator
    (conv padding modes, symmetric filter expansion, GroupNorm, erf-GELU, floor AvgPool,
    bicubic A=-0.75, curl + wall BCs, upwind/central stencil, CFL dt).
  * `ref_torch` -- the same algorithm expressed with stock ATen CPU ops, multi-threaded; this
    is the "port" that bench.py times as the CPU baseline (the reference itself is PyTorch).
"""
