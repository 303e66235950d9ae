#!/bin/bash
# Build everything that has to travel to the GPU box: the native library and the pre-built kernel
# specialisations (stale ones from older sources are removed first).
set -e
cd "$(dirname "$0")/.."
python -m genesis_forge_b200.build_native
rm -rf genesis_forge_b200/_spec
python tools/prebuild_specs.py
