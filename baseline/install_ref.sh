#!/bin/bash
# Install the UNMODIFIED reference (zhaoruiyang98/eftpipe, /root/reference) under baseline/_ref for the CPU arm of
# bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`) and for tests that drive the live reference.
# baseline/_ref is git-ignored (no reference source enters the history) but travels to the GPU box with gpurun.
#
# 1. pip install from a scratch copy (the build writes into the source tree; /root/reference is read-only; the 151 MB of
#    survey data are not part of the package), offline, without dependency resolution (cobaya is not installable here:
#    oracle/refshim/cobaya stands in for it).
# 2. The reference's pyproject.toml lists `packages = ["eftpipe"]` only, so the wheel omits the vendored sub-package
#    eftpipe/pybird (upstream installs in development mode, where this does not show).  The sub-package is completed from
#    the same source tree, byte for byte.
set -e
REF=${EFTPIPE_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
DEST="$HERE/_ref"
[ -d "$REF/eftpipe/pybird" ] || { echo "install_ref: no reference tree at $REF"; exit 0; }
TMP=$(mktemp -d)
mkdir -p "$TMP/src"
cp -r "$REF/eftpipe" "$REF/pyproject.toml" "$REF/README.md" "$REF/LICENSE" "$TMP/src/" 2>/dev/null || true
rm -rf "$DEST"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$DEST" "$TMP/src"
if [ ! -d "$DEST/eftpipe/pybird" ]; then
  cp -r "$REF/eftpipe/pybird" "$DEST/eftpipe/pybird"
fi
find "$DEST" -name __pycache__ -type d -prune -exec rm -rf {} +
rm -rf "$TMP"
# integrity: every installed file equals its source
( cd "$REF/eftpipe" && find . -type f \( -name '*.py' -o -name '*.yaml' \) -not -path '*/__pycache__/*' ) | while read -r f; do
  cmp -s "$REF/eftpipe/$f" "$DEST/eftpipe/$f" || { echo "install_ref: $f differs from the source"; exit 1; }
done
echo "install_ref: reference installed at $DEST ($(find "$DEST/eftpipe" -name '*.py' | wc -l) python files)"
