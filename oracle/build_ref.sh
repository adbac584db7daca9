#!/bin/bash
# TEST / MEASUREMENT INFRASTRUCTURE — not product code.
#
# Makes the UNMODIFIED reference runnable on the GPU box, where /root/reference does not exist:
# copies the reference's Python package and its grid files, as they are, from the read-only
# reference tree into oracle/_ref/ (git-ignored, so nothing of the reference enters the history;
# not gpurun-ignored, so it travels to the box with the snapshot).  bench.py --impl reference and
# the cpu_baseline leg of bench.py run it there through oracle/ref_shim.py
# (HELMPY_REFERENCE_ROOT=oracle/_ref) on the box's host cores (reference helmpy/core/helm.py:896).
#
# Usage: oracle/build_ref.sh [reference root, default /root/reference]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
if [ ! -d "$SRC/helmpy/core" ]; then
    echo "build_ref: no reference tree at $SRC (nothing to do)"; exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/data/cases"
cp -r "$SRC/helmpy" "$DST/helmpy"
cp "$SRC"/data/cases/*.xlsx "$DST/data/cases/"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
chmod -R u+w "$DST"
( cd "$SRC" && find helmpy data/cases -type f ! -name '*.pyc' -exec sha256sum {} + | sort -k2 ) > "$DST/SOURCE_SHA256"
echo "built $DST from $SRC ($(find "$DST/helmpy" -name '*.py' | wc -l) python files, $(ls "$DST/data/cases" | wc -l) grids)"
