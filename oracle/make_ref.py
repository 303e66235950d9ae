"""
TEST INFRASTRUCTURE (oracle).  Recipe for `oracle/_ref/`: a verbatim copy of the reference's Python
package, so that the UNMODIFIED reference can run where /root/reference does not exist (the GPU box).

    python -m oracle.make_ref            (build container; also run by __graft_entry__.build())

Copies /root/reference/genesis_forge (the package: managers, mdp, wrappers ...) and the gait_trainer
example's command manager (examples/gait_trainer/gait_command_manager.py, the user-level manager of
BASELINE config 3b) into oracle/_ref/, byte for byte, and writes oracle/_ref/MANIFEST.json with the
sha256 of every file so that a run can state exactly which sources it timed.  oracle/_ref/ is
git-ignored (no reference source enters the history) but not gpurun-ignored: it travels to the GPU box
like the built .so files do.  Nothing is compiled: the reference is pure Python.

Used by: bench.py --impl reference (`cpu_baseline.kind == "reference"`), oracle/shim.py as the
fallback location of the reference.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = "/root/reference"
TARGET = os.path.join(ROOT, "oracle", "_ref")
EXTRA_FILES = ["examples/gait_trainer/gait_command_manager.py"]


def make(source: str = SOURCE, target: str = TARGET) -> dict:
    package = os.path.join(source, "genesis_forge")
    if not os.path.isdir(package):
        raise FileNotFoundError(f"{package} not present")
    if os.path.isdir(target):
        shutil.rmtree(target)
    os.makedirs(target)
    shutil.copytree(package, os.path.join(target, "genesis_forge"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for rel in EXTRA_FILES:
        dst = os.path.join(target, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copy2(os.path.join(source, rel), dst)
    manifest = {}
    for base, _, files in os.walk(target):
        for name in sorted(files):
            path = os.path.join(base, name)
            with open(path, "rb") as f:
                manifest[os.path.relpath(path, target)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(target, "MANIFEST.json"), "w") as f:
        json.dump({"source": source, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


if __name__ == "__main__":
    files = make(*(sys.argv[1:3]))
    print(f"{TARGET}: {len(files)} files copied from {SOURCE}")
