"""``src.visymre.architectures.model`` of the stand-alone package: the names of the reference module
(``model.py:13-19``, ``:444-560``) that belong to the refinement path.

The reference's ``Model`` (a Lightning module: set-transformer encoder, transformer decoder,
beam loop -- SURVEY.md section 2.1, out of scope) is NOT rebuilt here.  To run the reference's
drivers, lay this package over a reference checkout with ``vision-sr_b200/overlay.py``: the checkout
keeps its own ``model.py`` and its ``Model.fitfunc2`` is patched to end in ``refine_hypotheses``.
"""
from .refine import (BINARY_NAMES, UNARY_NAMES, analyze_prefix_tree_context,  # noqa: F401
                     beam_constraint_mask, bfgs_wrapper, refine_hypotheses)


class Model:
    """Placeholder that fails loudly: the network lives in the reference checkout."""

    def __init__(self, *a, **k):
        raise NotImplementedError(
            "the ViSymRe network is the reference's own code (src/visymre/architectures/model.py); "
            "overlay this package on a reference checkout: python vision-sr_b200/overlay.py <checkout>")

    @classmethod
    def load_from_checkpoint(cls, *a, **k):
        return cls()

    @staticmethod
    def _analyze_prefix_tree_context(seq, arity_1_ids, arity_2_ids, transcendental_ids, pow_id, c_id, start_id=1):
        return analyze_prefix_tree_context(seq, arity_1_ids, arity_2_ids, transcendental_ids, pow_id, c_id, start_id)
