#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Vendor the UNMODIFIED reference (pure Python, no build step) into oracle/_ref/ so that
# it travels to the GPU box with the gpurun snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored):
#   * bench.py --impl reference  and the cpu_baseline leg time the reference's own modules on the host cores;
#   * tests/ use it as the live oracle when /root/reference is not mounted.
# Nothing is edited: a plain copy of /root/reference/src plus a MANIFEST of sha256 sums.  Never committed.
set -euo pipefail
cd "$(dirname "$0")"
SRC="${PDES_REFERENCE_ROOT:-/root/reference}"
if [[ ! -d "$SRC/src/models" ]]; then
  echo "oracle/make_ref.sh: $SRC/src not found (nothing to do; a prebuilt oracle/_ref is used if present)"
  exit 0
fi
rm -rf _ref
mkdir -p _ref
cp -r "$SRC/src" _ref/src
find _ref/src -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd _ref/src && find . -type f -name '*.py' | sort | xargs sha256sum ) > _ref/MANIFEST.sha256
echo "vendored $(wc -l < _ref/MANIFEST.sha256) reference files into oracle/_ref/src"
