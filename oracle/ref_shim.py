"""Import helper for the REAL reference (huggingface/diffusers at /root/reference) - build container only.

Used exclusively by oracle/make_golden.py and by CPU tests that pin the oracle; nothing that runs on the GPU
box may import this (the reference does not exist there).  The two huggingface_hub symbols are only used by
DiffusionPipeline.download (pipelines/pipeline_utils.py:1669-1685) and are missing from the installed hub.
"""
import os
import sys

REFERENCE_SRC = "/root/reference/src"


def available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "diffusers"))


def import_reference():
    if not available():
        raise ImportError("reference not present (expected on the GPU box)")
    import huggingface_hub
    import huggingface_hub.errors
    if not hasattr(huggingface_hub, "get_cached_repo_tree"):
        huggingface_hub.get_cached_repo_tree = lambda *a, **k: []
    if not hasattr(huggingface_hub.errors, "CachedRepoTreeNotFoundError"):
        class CachedRepoTreeNotFoundError(Exception):
            pass
        huggingface_hub.errors.CachedRepoTreeNotFoundError = CachedRepoTreeNotFoundError
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import diffusers
    return diffusers
