"""CPU oracle for the Student-t entropy bottleneck + GDN hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`domain_specific_image_compression_b200/`) imports this directory.  The only
legitimate callers are `tests/`, `__graft_entry__.smoke()` (as the checker) and
`bench.py`'s `cpu_baseline` / `--impl reference` legs (as the thing timed on the
host cores).  The product path is CUDA-only and raises when its extension is
missing; it never falls back to anything in here.

Every function cites the reference lines (relative to /root/reference) it
restates.  How each part is pinned:

* quantize / Student-t density NLL / Gaussian NLL / rate / GDN / IGDN / autograd
  gradients: pinned against the reference ITSELF, imported in the build
  container by `oracle/gen_golden.py`, outputs committed under `tests/golden/`.
* `pmf_to_uint16_cdf`: pinned against the reference's own function (imported
  with stub modules for its absent third-party imports).
* CDF-table *build* (support, Student-t / Gaussian CDF, PMF): the reference
  script cannot run as written and no test of the
  reference touches it => PARITY UNPINNED by the reference.  Pinned instead
  against scipy fp64 (`stdtr`, `ndtr`) and against a "repaired" run of the
  reference script (floor/ceil fixed, scipy-backed `.cdf`), see gen_golden.py.
* range coder: torchac 0.9.3 is not vendored and not installed => PARITY
  UNPINNED; the coder is ours on both sides, pinned by golden bytes + round trips.
* MS-SSIM (piq 0.8.0, absent): restated from its published algorithm => PARITY
  UNPINNED; cross-checked against an independent fp64 numpy/scipy evaluation.
"""
