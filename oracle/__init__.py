"""CPU oracle for the masked-window-attention / GDN / latent-rounding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this directory; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may.  The shipped path is the CUDA library behind `include/mwa_b200.h` and it fails
loudly when that library is missing.

Contents
  ref_ops.py         restatement of the reference algorithm (torch CPU tensors, fp32 or fp64),
                     each function citing the reference file:line it follows.
  live_reference.py  imports the *unmodified* reference modules from /root/reference through
                     the stand-in packages in shims/ (timm / compressai / tensorboardX are not
                     installed here).  Only available in the build container.
  make_golden.py     runs the live reference on seeded inputs and writes tests/golden/*.npz.

Pinning: the reference ships no tests or golden vectors.  ref_ops is pinned
against outputs of the reference itself run in the build container (tests/golden/, regenerated
by make_golden.py) and, when /root/reference is present, directly against the live modules
(tests/test_oracle_vs_reference.py).  The CompressAI boundary (likelihoods / bpp) is third-party,
absent and unpinned: "parity unpinned" there; it does not touch this path's outputs.
"""
