"""CPU oracle for the distllm embedding hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``distllm_b200/`` may import this package.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
(as the checker / the timed CPU reference, never as the product path).

Parity pin: the reference ships no golden vectors for this path, so the oracle
is pinned against outputs of the reference itself, run unmodified in the authoring container by
``oracle/make_golden.py`` (fixtures in ``tests/golden/``), and against HuggingFace ``BertModel``
(the third-party code that holds the reference's arithmetic; transformers is present on every box).
"""
