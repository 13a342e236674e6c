"""Import the UNMODIFIED reference hot-path modules (``min_tfs_client.tensors`` of zendesk/min-tfs-client).

Source of the code, in order: ``/root/reference`` (build container), else the archive
``baseline/_ref/min_tfs_client_reference.zip`` that ``baseline/stage_reference.py`` packed from it (what travels to
the GPU box).  The schema modules the reference generates with ``protoc`` at install time are this repo's
``tools/gen_pb2.py`` outputs (same descriptors, same protobuf runtime => same wire bytes).

This repo's drop-in package has the same import name, so the reference is loaded under the PRIVATE name
``_reference_min_tfs_client`` with an import hook that maps the reference's own absolute-free relative imports
(``from .types import DataType`` ...) onto that name.  Nothing of the product is imported here.

Only ``bench.py``'s CPU legs and tests use this module.
"""
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF_DIR = "/root/reference/tensor_serving_client/min_tfs_client"
ZIP = os.path.join(HERE, "_ref", "min_tfs_client_reference.zip")
NAME = "_reference_min_tfs_client"


def available():
    return os.path.isdir(REF_DIR) or os.path.exists(ZIP)


def load():
    """Returns (module ``tensors`` of the reference, description of where it came from)."""
    if NAME + ".tensors" in sys.modules:
        return sys.modules[NAME + ".tensors"], sys.modules[NAME].__origin__
    schema_root = os.path.join(REPO, "min-tfs-client_b200")
    if schema_root not in sys.path:
        sys.path.append(schema_root)      # for tensorflow.core.framework.*_pb2 only; the drop-in package itself is never imported
    if os.path.isdir(REF_DIR):
        path, origin = REF_DIR, "/root/reference (unmodified checkout)"
    elif os.path.exists(ZIP):
        path, origin = ZIP + "/min_tfs_client", "baseline/_ref/min_tfs_client_reference.zip (unmodified copy staged by baseline/stage_reference.py)"
    else:
        raise ImportError("the reference is neither at /root/reference nor staged in baseline/_ref (run baseline/stage_reference.py in the build container)")
    pkg = types.ModuleType(NAME)
    pkg.__path__ = [path]          # a namespace-style package rooted in the reference's directory (it has no __init__.py)
    pkg.__origin__ = origin
    sys.modules[NAME] = pkg
    tensors = importlib.import_module(NAME + ".tensors")
    return tensors, origin
