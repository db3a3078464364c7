"""Locates and imports the UNMODIFIED reference (huggingface/diffusers) for the reference arm of bench.py and for the
drop-in tests.  Not product code: diffusers_b200/ never imports this.

The reference is installed once, offline, with
    cp -r /root/reference /tmp/ref_src && python -m pip install --no-index --no-build-isolation --no-deps \
        --find-links /opt/wheelhouse --target baseline/_ref /tmp/ref_src
(`install()` below does exactly that).  baseline/_ref is git-ignored but NOT gpurun-ignored, so it travels to the GPU
box; on a box that has neither baseline/_ref nor /root/reference, `available()` is False and callers skip / report
"unavailable".

The image's huggingface_hub (1.14) lacks two symbols the reference imports at module level in
pipelines/pipeline_utils.py:32 and only uses in DiffusionPipeline.download (:1669-1685, never reached here: there is
no network and every component is passed in).  They are defined on the ENVIRONMENT's hub module, the reference's files
are not touched.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED = os.path.join(HERE, "_ref")
SOURCE_TREE = "/root/reference/src"


def location():
    if os.path.isdir(os.path.join(INSTALLED, "diffusers")):
        return INSTALLED
    if os.path.isdir(os.path.join(SOURCE_TREE, "diffusers")):
        return SOURCE_TREE
    return None


def available():
    return location() is not None


def install():
    """Build-container only (needs /root/reference and the offline wheelhouse)."""
    if os.path.isdir(os.path.join(INSTALLED, "diffusers")):
        return INSTALLED
    tmp = "/tmp/ref_src"
    if not os.path.isdir(tmp):
        shutil.copytree("/root/reference", tmp, symlinks=True)
    subprocess.check_call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
                           "/opt/wheelhouse", "--target", INSTALLED, tmp])
    return INSTALLED


def import_reference():
    loc = location()
    if loc is None:
        raise ImportError("the reference is neither installed under baseline/_ref nor present at /root/reference")
    import huggingface_hub
    import huggingface_hub.errors
    if not hasattr(huggingface_hub, "get_cached_repo_tree"):
        huggingface_hub.get_cached_repo_tree = lambda *a, **k: []
    if not hasattr(huggingface_hub.errors, "CachedRepoTreeNotFoundError"):
        class CachedRepoTreeNotFoundError(Exception):
            pass
        huggingface_hub.errors.CachedRepoTreeNotFoundError = CachedRepoTreeNotFoundError
    if loc not in sys.path:
        sys.path.insert(0, loc)
    import diffusers
    return diffusers
