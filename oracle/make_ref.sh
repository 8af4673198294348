#!/bin/sh
# Stages the UNMODIFIED reference package for the GPU box: copies /root/reference/src/SALib (pure Python, no
# build step) into oracle/_ref/SALib.  oracle/_ref/ is git-ignored (no reference source enters the history) but
# not gpurun-ignored, so it travels with the snapshot like the built .so files; /root/reference itself does not
# exist on the GPU box.  Used only by bench.py's CPU legs (`--impl reference`, `cpu_baseline`) and by tests that
# pin the oracle -- never by the product path.  Run by __graft_entry__.build() whenever /root/reference is present.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${SALIB_REFERENCE:-/root/reference}/src/SALib"
[ -d "$SRC" ] || { echo "make_ref: $SRC not found (nothing to do on the GPU box)"; exit 0; }
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC" "$HERE/_ref/SALib"
find "$HERE/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "${SALIB_REFERENCE:-/root/reference}" && { git rev-parse HEAD 2>/dev/null || echo "no-git"; } ) > "$HERE/_ref/SOURCE_REVISION" || true
( cd "$HERE/_ref" && find SALib -name '*.py' | sort | xargs sha256sum | sha256sum | cut -d' ' -f1 ) > "$HERE/_ref/TREE_SHA256"
echo "make_ref: staged $(find "$HERE/_ref/SALib" -name '*.py' | wc -l) files, tree sha256 $(cat "$HERE/_ref/TREE_SHA256")"
