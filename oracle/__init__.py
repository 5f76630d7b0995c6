"""CPU oracle for the self-paced SupCon hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package may import this; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference arm do,
and only as the checker or the timed CPU baseline.

Two restatements of ``contrastyou/losses/contrast_loss3.py`` (reference):

* :mod:`oracle.closed_form` -- fp64 numpy, blockwise closed form (never builds
  N x N for large N), loss + ratio + analytic gradient.
* :mod:`oracle.dense_port`  -- fp32 torch-CPU port that follows the reference's
  dense N x N op sequence (same cost structure; used as the CPU baseline).

Parity pin: the reference ships no golden vectors for this path (SURVEY.md
section 8c).  Both restatements are pinned against outputs of the *unmodified*
reference module executed in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
"""
