"""visymre (B200 build): the post-decode refinement path of ViSymRe.

Only the hot path of the reference is rebuilt here -- constant fitting of beam
candidates (reference ``src/visymre/architectures/bfgs.py`` and the "BFGS Parallel
Part" of ``Model.fitfunc2``, ``src/visymre/architectures/model.py:444-520``).  The
package keeps the reference's module paths (``src.visymre.architectures.bfgs`` ...)
so its ``scripts/*_test.py`` drivers import it unchanged.  All numeric work runs in
the CUDA extension under ``vision-sr_b200/csrc``; there is no CPU fallback.
"""
