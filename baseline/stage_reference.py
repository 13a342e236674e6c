#!/usr/bin/env python
"""Stage the UNMODIFIED reference modules of the hot path for the GPU box.

``/root/reference`` exists only in the build container; ``gpurun`` ships ``/root/repo`` (git-ignored files
included).  This script packs the reference's three hot-path modules

    tensor_serving_client/min_tfs_client/{tensors,types,constants}.py

byte for byte into ``baseline/_ref/min_tfs_client_reference.zip`` (git-ignored: no reference source enters the
history) together with a manifest of their SHA-256 digests.  ``baseline/ref_loader.py`` imports them from
``/root/reference`` when it is there, else from the zip (zipimport), over this repo's generated ``*_pb2`` schema
modules - the same arrangement ``tests/golden/make_golden.py`` uses to produce the golden vectors.
``bench.py --impl reference`` and the ``cpu_baseline`` leg then time that code (``kind: "reference"``).

    python baseline/stage_reference.py        # idempotent; exits 0 and says so when the reference is absent
"""
import hashlib
import json
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PKG = "/root/reference/tensor_serving_client/min_tfs_client"
FILES = ("tensors.py", "types.py", "constants.py")
OUT_DIR = os.path.join(HERE, "_ref")
ZIP = os.path.join(OUT_DIR, "min_tfs_client_reference.zip")


def stage():
    if not os.path.isdir(REF_PKG):
        print("reference not present (not the build container): nothing staged")
        return False
    os.makedirs(OUT_DIR, exist_ok=True)
    manifest = {}
    blobs = {}
    for f in FILES:
        with open(os.path.join(REF_PKG, f), "rb") as fh:
            blobs[f] = fh.read()
        manifest[f] = hashlib.sha256(blobs[f]).hexdigest()
    if os.path.exists(ZIP):
        try:
            with zipfile.ZipFile(ZIP) as z:
                if json.loads(z.read("MANIFEST.json")) == manifest:
                    return True
        except Exception:  # noqa: BLE001 - rewrite a damaged archive
            pass
    tmp = ZIP + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            z.writestr(zipfile.ZipInfo("min_tfs_client/" + f, date_time=(2020, 1, 1, 0, 0, 0)), blobs[f])
        z.writestr(zipfile.ZipInfo("MANIFEST.json", date_time=(2020, 1, 1, 0, 0, 0)), json.dumps(manifest, sort_keys=True))
    os.replace(tmp, ZIP)
    print("staged", ZIP, manifest)
    return True


if __name__ == "__main__":
    stage()
    sys.exit(0)
