"""CPU oracle for the distllm embedding hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``distllm_b200/`` may import this package.  Allowed users: ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
(as the checker / the timed CPU reference, never as the product path).

Modules: ``bert`` / ``esm`` / ``mistral`` (the three forward passes the reference's encoders trigger in
HuggingFace transformers), ``pooling`` (mean / last-token poolers and the batch loop), ``semantic``
(adjacent-buffer distances and the percentile split), ``search`` (exact inner-product top-k, i.e.
faiss.IndexFlatIP as the reference's query path calls it), ``make_golden`` (fixture generator).

Parity pin: the reference ships no golden vectors for this path, so the oracle
is pinned against outputs of the reference itself, run unmodified in the authoring container by
``oracle/make_golden.py`` (fixtures in ``tests/golden/``), and against HuggingFace ``BertModel`` /
``EsmModel`` / ``MistralModel`` (the third-party code that holds the reference's arithmetic;
transformers is present on every box).  ``search`` is the one exception: faiss is absent from this
image, so it restates IndexFlatIP's published behaviour and is pinned only by its own brute-force
check -- "parity unpinned" for that module.
"""
