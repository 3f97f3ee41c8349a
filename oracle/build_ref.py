"""Recipe for oracle/_ref: the UNMODIFIED reference hot path, byte-compiled where it lies.

The reference is pure Python (no build system).  This script compiles the few modules its
``bfgs()`` needs -- straight from ``/root/reference/src/visymre`` -- to SOURCELESS ``.pyc`` files
under ``oracle/_ref/src/visymre/...`` (``py_compile``; no reference source text is copied into the
repo) and extracts the pickled vocabulary record from ``scripts/weights/meta/metadata.h5`` (bytes
2048..4974).  ``oracle/_ref/`` is git-ignored but travels to the GPU box, where
``/root/reference`` does not exist: there ``bench.py --impl reference --config 1`` times this
as-is reference (``cpu_baseline.kind = "reference"``) and ``oracle/ref_harness.load()`` finds it.

Run:  python oracle/build_ref.py      (``__graft_entry__.build()`` does, when /root/reference exists)
TEST INFRASTRUCTURE: nothing under vision-sr_b200/ imports oracle/.
"""
import os
import py_compile
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
# what `import src.visymre.architectures.bfgs` / `.model` pull in, nothing else
MODULES = [
    "src/__init__.py", "src/visymre/__init__.py", "src/visymre/dclasses.py", "src/visymre/utils.py",
    "src/visymre/architectures/__init__.py", "src/visymre/architectures/bfgs.py",
    "src/visymre/architectures/data.py", "src/visymre/architectures/model.py",
    "src/visymre/architectures/beam_search.py", "src/visymre/architectures/MultimodalEncoder.py",
    "src/visymre/dataset/__init__.py", "src/visymre/dataset/generator.py",
    "src/visymre/dataset/sympy_utils.py", "src/visymre/dataset/data_utils.py",
]


def build(ref=REF, out=OUT):
    if not os.path.isdir(ref):
        raise FileNotFoundError(f"{ref} is not here: oracle/_ref can only be built in the build container")
    if os.path.isdir(out):
        shutil.rmtree(out)
    n = 0
    for rel in MODULES:
        src = os.path.join(ref, rel)
        dst = os.path.join(out, rel + "c")          # module.pyc beside where module.py would be
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(src):
            if rel.endswith("__init__.py"):         # namespace package in the reference: an empty module
                tmp = dst[:-1] + ".empty"
                open(tmp, "w").close()
                py_compile.compile(tmp, cfile=dst, doraise=True, dfile=rel)
                os.remove(tmp)
                n += 1
            continue
        py_compile.compile(src, cfile=dst, doraise=True, dfile=src)
        n += 1
    raw = open(os.path.join(ref, "scripts/weights/meta/metadata.h5"), "rb").read()
    with open(os.path.join(out, "metadata.pkl"), "wb") as fh:
        fh.write(raw[2048:2048 + 2926])
    with open(os.path.join(out, "PYTHON"), "w") as fh:
        fh.write(f"{sys.version_info.major}.{sys.version_info.minor}\n")   # .pyc files are per minor version
    return n


if __name__ == "__main__":
    print(f"oracle/_ref: {build()} modules byte-compiled from {REF}")
