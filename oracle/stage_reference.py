"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the five UNMODIFIED reference hot-path files under oracle/_ref/ (git-ignored).

    python -m oracle.stage_reference        (also run by __graft_entry__.build() whenever /root/reference is mounted)

Why: the reference is pure Python and /root/reference does not exist on the GPU box, so `bench.py --impl reference` could only time
the restatement (kind "port") there.  oracle/_ref/ is git-ignored but travels with the snapshot, exactly like the built .so files:
with the files staged, the reference arm and `cpu_baseline` run the reference's OWN `TripletE2ENet.step` + backward (kind
"reference") through oracle/ref_shim.py (third-party imports stubbed as in SURVEY.md Appendix A).  The copies are byte-for-byte
(sha256 recorded in MANIFEST.json and re-checked on load); nothing under oracle/_ref/ is ever committed, imported by the product
package or used as anything but the checker / baseline.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("IB200_REFERENCE_SOURCE", "/root/reference")
FILES = ("intrepppid/utils/weightdrop.py", "intrepppid/utils/embedding_do.py", "intrepppid/encoders/awd_lstm.py",
         "intrepppid/classifier/head/mlp.py", "intrepppid/e2e/e2e_triplet.py")


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def stage(verbose: bool = False) -> bool:
    """Copy the files if the reference is mounted; returns True when oracle/_ref/ holds a verified set afterwards."""
    if os.path.isfile(os.path.join(SOURCE, FILES[0])):
        manifest = {"source": SOURCE, "files": {}}
        for rel in FILES:
            dst = os.path.join(DEST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(SOURCE, rel), dst)
            manifest["files"][rel] = _sha(dst)
        with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
            json.dump(manifest, fh, indent=1)
        if verbose:
            print(f"staged {len(FILES)} reference files under {DEST}")
    return verify()


def verify() -> bool:
    path = os.path.join(DEST, "MANIFEST.json")
    if not os.path.isfile(path):
        return False
    with open(path) as fh:
        manifest = json.load(fh)
    return all(os.path.isfile(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == sha
               for rel, sha in manifest["files"].items()) and set(manifest["files"]) == set(FILES)


if __name__ == "__main__":
    print("verified" if stage(verbose=True) else "reference not available: nothing staged")
