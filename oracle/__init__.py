"""TEST INFRASTRUCTURE — not product code.

``oracle`` holds the CPU restatement of the FCMA correlation hot path of
brainiak/brainiak @ 123f6e1 (``fcma_oracle.c`` + ``fcma_oracle.py``) and the loader for the
unmodified reference built from source (``reference.py`` -> ``oracle/_ref``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``brainiak_b200`` never does: its product
path fails loudly when the CUDA library is missing instead of falling back to anything here.

Parity status: PINNED (see the header of ``fcma_oracle.c``).
"""
