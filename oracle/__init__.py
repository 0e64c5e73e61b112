"""CPU oracle for the LARIS per-cell LR scoring path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code: it may be imported by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs, and by nothing else.  ``laris_b200`` never imports it.

* ``oracle.restate``   - numpy/scipy/sklearn restatement of the reference's
  algorithm for this path; each function cites the reference file:line it
  follows.  The reference's own arithmetic lives in third-party compiled
  routines (scikit-learn >=0.24, scipy >=1.7, numpy legacy MT19937;
  reference pyproject.toml:37-47) which ARE importable here and are called
  the same way the reference calls them.
* ``oracle.rng``       - numpy Philox4x32-10 + Feistel permutation, the
  bit-exact CPU twin of the device counter-based RNG.
* ``oracle.ref_loader``- loads the UNMODIFIED reference (``/root/reference``)
  under stub modules; only usable in the build container.  Used by
  ``oracle/make_golden.py`` to pin the restatement and write ``tests/golden``.

Parity status: the reference ships no tests, golden vectors or fixtures
(SURVEY.md §4, §8c).  The restatement is pinned against outputs of the
reference itself run in the build container (``tests/golden/*.npz``, generated
by ``oracle/make_golden.py``) for every function that runs unmodified there;
``cosg.indexByGene`` / ``cosg.iqrLogNormalize`` / ``cosg.cosg`` (package absent,
source not vendored) are **parity unpinned**.
"""
