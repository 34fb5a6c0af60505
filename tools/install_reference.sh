#!/bin/bash
# Installs the UNMODIFIED reference (zjykzj/YOLOv4, /root/reference) into baseline/_ref/ (git-ignored; it travels to the GPU
# box with the gpurun snapshot).  The reference ships no setup.py, so the install runs from a copy under /tmp that gets a
# five-line one; no reference source is changed and none enters the repository's history.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference}
[ -d "$SRC/yolo" ] || { echo "no reference at $SRC"; exit 1; }
TMP=$(mktemp -d)
cp -r "$SRC" "$TMP/ref"
cat > "$TMP/ref/setup.py" <<'PY'
from setuptools import setup, find_namespace_packages
setup(name="zjykzj-yolov4-reference", version="0", packages=find_namespace_packages(include=["yolo", "yolo.*", "darknet", "darknet.*"]),
      py_modules=[], data_files=[("config", ["config/yolov4_default.cfg", "config/yolov4_Tianxiaomo.cfg"])])
PY
[ -f "$TMP/ref/darknet/__init__.py" ] || touch "$TMP/ref/darknet/__init__.py"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP/ref" 2>&1 | tail -3
rm -rf "$TMP"
ls "$ROOT/baseline/_ref"
