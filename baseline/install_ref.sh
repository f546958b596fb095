#!/usr/bin/env bash
# Installs the UNMODIFIED reference (astra-vision/LatteCLIP, /root/reference) into baseline/_ref
# (git-ignored; it travels to the GPU box with the repository snapshot).  /root/reference is
# read-only and setup.py writes build files, so the install runs from a copy under /tmp;
# dependency resolution is skipped (--no-deps: torch is already in the image, the wheelhouse has
# no torch wheel).  bench.py --impl reference loads baseline/_ref/open_clip/loss.py from here.
set -euo pipefail
cd "$(dirname "$0")/.."
rm -rf /tmp/latteclip_ref_copy baseline/_ref
cp -r /root/reference /tmp/latteclip_ref_copy
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref /tmp/latteclip_ref_copy
rm -rf /tmp/latteclip_ref_copy
ls baseline/_ref
