"""Recipe for ``oracle/_ref/``: byte-compile the UNMODIFIED reference modules where they lie.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (onolab-tmu/overiva) is three pure-Python files; ``/root/reference``
exists in the build container but not on the GPU box, and reference SOURCES must never be copied into this repository.
So, exactly like a C reference would be compiled into ``oracle/_ref/*.so``, the Python reference is compiled -- from
the sources where they lie, nothing is copied -- into CPython bytecode files ``oracle/_ref/<module>.refbc`` (a ``.pyc`` image under
another extension: snapshot tools commonly drop ``*.pyc``; build outputs: git-ignored, they travel to the GPU box with
the snapshot like the built ``.so``).  ``oracle/reference_shim.py`` imports
the bytecode when the source tree is absent, so that ``bench.py --impl reference`` and ``cpu_baseline`` time the REAL
reference implementation (``kind: "reference"``) on the box's host cores.

    python -m oracle.build_ref          # run by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
MODULES = ("overiva", "auxiva_pca", "ive")  # the three files on the hot path


def build(reference_dir="/root/reference") -> bool:
    """Returns True when oracle/_ref/ holds bytecode of all three modules for this interpreter."""
    if not os.path.isfile(os.path.join(reference_dir, "overiva.py")):
        return all(os.path.isfile(os.path.join(REF_OUT, m + ".refbc")) for m in MODULES)
    os.makedirs(REF_OUT, exist_ok=True)
    for m in MODULES:
        py_compile.compile(os.path.join(reference_dir, m + ".py"), cfile=os.path.join(REF_OUT, m + ".refbc"),
                           dfile="<reference>/%s.py" % m, doraise=True, optimize=0,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(REF_OUT, "BUILT_WITH"), "w") as f:
        f.write("python %d.%d (magic %s) from %s\n" % (sys.version_info[0], sys.version_info[1],
                                                      __import__("importlib.util").util.MAGIC_NUMBER.hex(),
                                                      reference_dir))
    return True


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref:", "built" if ok else "reference tree absent, nothing built")
    sys.exit(0 if ok else 1)
