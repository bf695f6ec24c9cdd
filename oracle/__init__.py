"""Oracle package: CPU restatements of the reference algorithm (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
legs may import this package.  The product (``scico_b200``) never does.
"""
