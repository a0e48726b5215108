#!/bin/bash
# One-off, offline: install the UNMODIFIED reference package into baseline/_ref (git-ignored, travels to the GPU box
# with gpurun snapshots) so that bench.py's reference-torch leg and tools/run_config5.py can import it there.
#   pip install --no-index --no-build-isolation --no-deps --target baseline/_ref  <copy of source/SwarmACB_isaac>
# (the reference tree is read-only and setup.py writes build/ next to itself, hence the /tmp copy; --no-deps because
# its only declared dependency, psutil, is already in the image).  The training YAMLs and scripts/manual_control.py
# live outside the python package; they are copied next to it as data (never into git).
set -e
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DST=$ROOT/baseline/_ref
rm -rf /tmp/_swarm_ref_copy "$DST"
cp -r "$REF/source/SwarmACB_isaac" /tmp/_swarm_ref_copy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DST" /tmp/_swarm_ref_copy
mkdir -p "$DST/_extras"
cp -r "$REF/configs" "$DST/_extras/configs"
cp "$REF/scripts/manual_control.py" "$DST/_extras/manual_control.py"
rm -rf /tmp/_swarm_ref_copy
du -sh "$DST"
