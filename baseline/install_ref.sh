#!/bin/bash
# Install the UNMODIFIED reference (the diffusers 0.27 fork of val-iisc/Reflecting-Reality) into baseline/_ref with pip --target.
# baseline/_ref is git-ignored (no reference sources in history) and NOT gpurun-ignored (it travels to the GPU box), so
# `bench.py --impl reference`, `cpu_baseline` and `gpu_eager_baseline` run the reference's own StableDiffusionBrushNetPipeline there.
# /root/reference is read-only and setuptools writes build/ and *.egg-info into the source tree, hence the /tmp copy.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference/MirrorFusion}"
[ -d "$SRC/src/diffusers" ] || { echo "install_ref: $SRC is not the reference tree" >&2; exit 1; }
TMP="$(mktemp -d /tmp/mfref.XXXXXX)"
cp -r "$SRC/." "$TMP/"
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP"
rm -rf "$TMP"
find "$HERE/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "install_ref: $(du -sh "$HERE/_ref" | cut -f1) in $HERE/_ref"
