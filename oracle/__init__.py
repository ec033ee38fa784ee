"""CPU oracle for the DI-Fusion per-frame map hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a CPU restatement (torch-CPU fp32 / numpy) of the
reference algorithms, each function citing the reference file:line it follows.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Nothing under ``nerf-fusion_b200/`` imports it; the product path has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * map / network / tracker-SDF functions: PINNED against the reference's own Python
    (``/root/reference/system/map.py``, ``network/*``, ``system/tracker.py``) imported in the build
    container through ``oracle/ref_shims.py``; golden vectors in ``tests/golden/`` were produced by
    ``oracle/make_golden.py`` from that run; ``tests/test_oracle_cpu.py`` re-checks the oracle against them
    on every CPU run (ids / masks / counts / preprocessing bit-exact, H, g, poses, cubes, triangles).
  * pose algebra: PINNED against ``utils/motion_util.py`` itself (``tests/test_motion_cpu.py``, build container).
  * frame ingest: PINNED against the reference's own ``ICLNUIMSequence`` (``oracle/make_dataset_golden.py``,
    ``tests/golden/dataset_golden.npz``, ``tests/test_dataset_cpu.py``).
  * CUDA-extension ops (``system/ext/*``): the reference ships no tests or golden vectors and its
    kernels cannot run in the build container (no GPU).  They are restated here from the sources
    and pinned on the GPU box against the reference's own extensions when ``oracle/_ref/ext_build``
    (built from /root/reference by ``oracle/build_ref_ext.py``) loads; otherwise PARITY UNPINNED for
    those ops.
"""
