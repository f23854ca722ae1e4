#!/usr/bin/env bash
# Build the UNMODIFIED reference (r9y9/nnmnkwii, Cython/bandmat CPU path) into oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (never committed) but travels to the GPU
# box with the gpurun snapshot.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
# --impl reference legs may import it.  Nothing here copies reference sources into the tracked repo:
# the reference tree is copied to a scratch dir under /tmp (it is read-only and its build writes
# into the source tree), installed with pip --target, and the scratch dir is removed.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NNK_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/nnmnkwii" ]; then
  echo "build_ref: $REF not present (GPU box?) - using prebuilt oracle/_ref if any" >&2
  exit 0
fi
if [ -f "$OUT/.built" ] && [ "${1:-}" != "--force" ]; then
  echo "build_ref: $OUT already built"; exit 0
fi
TMP="$(mktemp -d /tmp/nnk_ref_build.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF/nnmnkwii" "$REF/setup.py" "$REF/README.md" "$REF/MANIFEST.in" "$TMP/" 2>/dev/null || true
[ -f "$REF/pyproject.toml" ] && cp "$REF/pyproject.toml" "$TMP/"
rm -rf "$OUT"; mkdir -p "$OUT"
( cd "$TMP" && python -m pip install --no-index --no-build-isolation --no-deps \
      --find-links /opt/wheelhouse --target "$OUT" . ) > "$TMP/pip.log" 2>&1 || {
  echo "build_ref: pip install failed, falling back to build_ext --inplace" >&2
  tail -20 "$TMP/pip.log" >&2
  ( cd "$TMP" && echo "__version__ = '0.1.3'" > nnmnkwii/version.py && python setup.py build_ext --inplace ) > "$TMP/build.log" 2>&1
  cp -r "$TMP/nnmnkwii" "$OUT/nnmnkwii"
  find "$OUT" -name '*.pyx' -delete -o -name '*.c' -delete
}
python - <<PY
import sys; sys.path.insert(0, "$OUT")
import nnmnkwii.paramgen as G, numpy as np
w=[(0,0,np.array([1.0])),(1,1,np.array([-0.5,0,0.5])),(1,1,np.array([1.0,-2.0,1.0]))]
y=G.mlpg(np.random.rand(10,6),np.random.rand(10,6)+.1,w); assert y.shape==(10,2)
print("build_ref: reference import OK ->", G.__file__)
PY
touch "$OUT/.built"
