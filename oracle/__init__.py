"""CPU oracle for the Point-NeRF per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithm of the reference's hot path
(SHUzhekiNg/pointnerf2studio, `pointnerf-original` plugin).  It exists so the CUDA
path in ``pointnerf2studio_b200`` can be checked; it is never the thing that is
shipped or measured.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Pinning status (see DESIGN.md "Oracle"):
  * field / compositing stages (rows W, E, M1, A, M2, D, C of SURVEY.md section 8a) are
    PINNED: ``tests/golden/make_golden.py`` executes the reference's own
    ``PointAggregator`` + ``ray_march`` + ``positional_encoding`` +
    ``near_far_linear_ray_generation`` (imported from /root/reference with three
    stub modules) with the shipped trained ``aggregator.*`` weights and the oracle
    reproduces those tensors (``tests/test_oracle_golden.py``).
  * neighbour query (rows G1, G2, Q): the reference implementation is CUDA-only
    and racy; the oracle is a deterministic restatement (two independent ones: numpy
    in ``grid_query.py`` and C in ``query_oracle.c``).  It is pinned against the
    unmodified reference extension compiled into ``oracle/_ref`` on the GPU box
    (``tests/test_gpu_reference_querier.py``); on a CPU-only box that pin cannot
    run, and there the two restatements only pin each other.
  * plugin-mode deltas (ReLU density, no confidence in the weight, Nerfstudio
    RGBRenderer background) follow nerfstudio (>=0.3.0, not vendored, not
    installable here): parity UNPINNED for those three one-line deltas.
"""
