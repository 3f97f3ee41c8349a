"""``src.visymre.utils`` of the stand-alone package: the one helper of the reference module the
refinement path needs -- ``load_metadata_hdf5`` (``utils.py:254-261``), which the drivers call
through ``scripts/visymre_utils.py:75`` to get the vocabulary record (``DatasetDetails``).

The reference reads the pickle out of the HDF5 dataset ``"other"`` with h5py.  Without h5py the
same bytes are found by their position in the shipped file (the dataset is stored contiguously;
for ``scripts/weights/meta/metadata.h5`` it starts at byte 2048): the pickle stream is located by
its protocol header and unpickled as far as it goes.
"""
import os
import pickle


def load_metadata_hdf5(path_folder):
    path = os.path.join(path_folder, "metadata.h5")
    try:
        import h5py
        import numpy as np
        with h5py.File(path, "r") as f:
            return pickle.loads(np.array(f["other"]).tobytes())
    except ImportError:
        pass
    raw = open(path, "rb").read()
    for start in (2048,) + tuple(i for i in range(0, len(raw) - 2) if raw[i] == 0x80 and raw[i + 1] in (2, 3, 4, 5)):
        try:
            return pickle.loads(raw[start:])     # unpickling stops at the STOP opcode
        except Exception:  # noqa: BLE001
            continue
    raise ValueError(f"no pickled metadata record found in {path}")
