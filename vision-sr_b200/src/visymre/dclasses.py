"""Config and metadata records of the refinement path.

Field names follow the reference (``src/visymre/dclasses.py:65-104``) so that
``scripts/visymre_utils.py:77-94`` can keep building ``BFGSParams``/``FitParams``
and so that the pickled ``DatasetDetails`` inside ``scripts/weights/meta/metadata.h5``
(written with the module path ``src.visymre.dclasses``) still unpickles.
Training-only records (``Equation``, ``NNEquation``, ``DataModuleParams``) are kept
as plain containers for the same reason; nothing on the hot path reads them.
"""
from dataclasses import dataclass, field
from typing import Optional


@dataclass
class GeneratorDetails:
    max_len: int = 0
    operators: str = ""
    max_ops: int = 0
    rewrite_functions: str = ""
    variables: list = field(default_factory=list)
    eos_index: int = 1
    pad_index: int = 0


@dataclass
class DatasetDetails:
    config: dict
    total_coefficients: list
    total_variables: list
    word2id: dict
    id2word: dict
    una_ops: list
    bin_ops: list
    rewrite_functions: list
    total_number_of_eqs: int = 0
    eqs_per_hdf: int = 0
    generator_details: Optional[GeneratorDetails] = None
    unique_index: Optional[set] = None


@dataclass
class Equation:
    expr: str = ""
    eq_sympy_prefix: list = field(default_factory=list)
    coeff_dict: dict = field(default_factory=dict)
    variables: list = field(default_factory=list)
    support: Optional[tuple] = None
    tokenized: Optional[list] = None
    tokenized_constant: Optional[list] = None
    valid: bool = True
    number_of_points: Optional[int] = None
    tokenized2: Optional[list] = None


@dataclass
class BFGSParams:
    # reference dclasses.py:83-91 (note idx_remove defaults to True here but the
    # shipped scripts/config.yaml:123 sets it False)
    activated: bool = True
    n_restarts: int = 10
    add_coefficients_if_not_existing: bool = False
    normalization_o: bool = False
    idx_remove: bool = True
    normalization_type: str = "MSE"
    stop_time: float = 1e9


@dataclass
class FitParams:
    word2id: dict
    id2word: dict
    total_coefficients: list
    total_variables: list
    rewrite_functions: list
    una_ops: Optional[list] = None
    bin_ops: Optional[list] = None
    bfgs: BFGSParams = field(default_factory=BFGSParams)
    beam_size: int = 2
    device: Optional[str] = None
