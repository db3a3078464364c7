"""Import helper for the REAL reference (huggingface/diffusers), used by oracle/make_golden.py and by tests that pin
the oracle or drive the shells from the unmodified reference.  The work is done by baseline/ref_env.py (installed copy
under baseline/_ref first, /root/reference/src in the build container otherwise)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import ref_env  # noqa: E402

REFERENCE_SRC = ref_env.SOURCE_TREE


def available():
    return ref_env.available()


def import_reference():
    return ref_env.import_reference()
