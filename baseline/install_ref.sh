#!/bin/bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored; travels to the GPU box with gpurun) for the
# reference arm of bench.py.  Needs /root/reference, i.e. runs in the build container only (__graft_entry__.build()).
#   1. the contract's offline pip install, from a copy (the source tree is read-only), with --no-deps: the image has
#      no wheels for xarray / netCDF4 / tables / pymc / func_timeout / tomlkit / dask (oracle/standins provides the
#      three the hot path imports);
#   2. the reference's pyproject lists packages = ["attrici"] only, so a non-editable install omits the sub-packages
#      attrici.commands / attrici.estimation / attrici.vendored: they are copied next to it, unmodified.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${1:-/root/reference}
[ -d "$REF/attrici" ] || { echo "no reference at $REF"; exit 0; }
TMP=$(mktemp -d)
cp -r "$REF" "$TMP/ref"
rm -rf "$ROOT/baseline/_ref"
SETUPTOOLS_SCM_PRETEND_VERSION=0.0.0 python -m pip install -q --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$ROOT/baseline/_ref" "$TMP/ref" || mkdir -p "$ROOT/baseline/_ref/attrici"
for sub in commands estimation vendored; do
    [ -d "$ROOT/baseline/_ref/attrici/$sub" ] || cp -r "$REF/attrici/$sub" "$ROOT/baseline/_ref/attrici/$sub"
done
for f in "$REF"/attrici/*.py; do
    [ -f "$ROOT/baseline/_ref/attrici/$(basename "$f")" ] || cp "$f" "$ROOT/baseline/_ref/attrici/"
done
find "$ROOT/baseline/_ref" -name __pycache__ -prune -exec rm -rf {} +
rm -rf "$TMP"
diff -rq "$REF/attrici" "$ROOT/baseline/_ref/attrici" && echo "baseline/_ref/attrici identical to $REF/attrici"
