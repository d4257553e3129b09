#!/bin/bash
# Install the UNMODIFIED reference (pydata/sparse, /root/reference) into baseline/_ref/ -- git-ignored, but shipped to
# the GPU box by gpurun -- so that bench.py's reference arm times the reference's own numba path
# (sparse.tensordot -> _dot_csr_ndarray, numba_backend/_common.py:95,720-755) on the box's host cores.
#
#   bash tools/make_ref.sh            (authoring container only: needs /root/reference; no network)
#
# /root/reference is read-only and pip wants to write build files, so the install runs from a copy under /tmp.
# setuptools-scm is not in the image, hence the wheel lacks the generated sparse/_version.py that
# sparse/__init__.py:5 imports; the two-line stub below is that generated file, nothing else is touched.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
DEST="$ROOT/baseline/_ref"
[ -d "$REF/sparse" ] || { echo "make_ref: $REF/sparse not found (reference only exists in the authoring container)"; exit 0; }
TMP="$(mktemp -d /tmp/ref_src.XXXXXX)"
cp -r "$REF/." "$TMP/"
rm -rf "$DEST"
mkdir -p "$DEST"
python -m pip install -q --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$DEST" "$TMP" \
  || { echo "make_ref: pip install failed; copying the package directory instead"; cp -r "$REF/sparse" "$DEST/sparse"; }
[ -f "$DEST/sparse/_version.py" ] || printf '__version__ = "0.0.0+ref"\n__version_tuple__ = (0, 0, 0)\n' > "$DEST/sparse/_version.py"
rm -rf "$TMP" 2>/dev/null || true
( cd /tmp && PYTHONPATH="$DEST" python -c "import sparse; print('baseline/_ref: sparse', sparse.__version__, 'backend', sparse._BACKEND)" )
