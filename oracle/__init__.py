"""CPU oracle for the FragNet GAT2 hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker (or as the CPU baseline
being timed) -- never as the thing the product computes with.

Contents
--------
``shims.py``        pure-torch stand-ins for the third-party ops the reference calls
                    (``torch_scatter.scatter_add / scatter_softmax``,
                    ``torch_geometric.utils.add_self_loops``), restating their published
                    semantics.  torch_scatter is un-pinned in the
                    reference (``install_gpu.sh:5``); torch_geometric is pinned ==2.6.1
                    (``requirements.txt:20``).  Neither is installed in this image.
``ref_import.py``   loads the UNMODIFIED reference files from ``/root/reference`` with the
                    shims injected.  Works only where the reference checkout exists (the
                    build container); it never travels to the GPU box.
``gat2_oracle.py``  a plain-PyTorch CPU restatement of the reference algorithm, op by op,
                    each function citing the reference file:line it follows.  This is the
                    oracle the GPU parity tests run against.
``collate_oracle.py`` restatement of the reference's Python-loop batch assembly.

Pinning status: **parity unpinned by the reference itself** -- the reference ships no
tests, golden vectors or loadable checkpoints for this path.  The
restatement is therefore pinned against outputs of the reference code itself, run in the
build container: ``tests/test_oracle_vs_reference.py`` compares ``gat2_oracle`` with the
unmodified reference (via ``ref_import``) whenever ``/root/reference`` is present, and
``tests/golden/make_golden.py`` stores reference outputs as fixtures that travel.
"""
