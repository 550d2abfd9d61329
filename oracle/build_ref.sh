#!/usr/bin/env bash
# Builds the UNMODIFIED reference Chamfer extension (extensions/chamfer_dist/{chamfer.cu,chamfer_cuda.cpp},
# setup.py:11-19) for sm_100a into oracle/_ref/ so the GPU parity tests can run the reference's own kernel
# beside ours on the B200 box.  Test infrastructure only: nothing in the product path imports it.
# Sources are compiled from a scratch copy (setup.py writes build/ next to itself and /root/reference is
# read-only); no reference source is copied into this repo.  oracle/_ref/ is git-ignored.
set -euo pipefail
REF=${REF:-/root/reference}
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/extensions/chamfer_dist" ]; then
  echo "build_ref: $REF not present (GPU box) - using prebuilt files in $OUT if any"; exit 0
fi
mkdir -p "$OUT"
if ls "$OUT"/chamfer*.so >/dev/null 2>&1; then
  echo "build_ref: $OUT/chamfer*.so already built"
else
  SCRATCH="$(mktemp -d /tmp/ref_chamfer.XXXXXX)"
  cp "$REF"/extensions/chamfer_dist/{chamfer.cu,chamfer_cuda.cpp,setup.py} "$SCRATCH"/
  ( cd "$SCRATCH" && TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=4 python setup.py -q build_ext --build-lib "$OUT" )
  rm -rf "$SCRATCH"
fi
# the auction-EMD extension (extensions/emd/{emd.cpp,emd_cuda.cu}, setup.py), same recipe
if ls "$OUT"/emd*.so >/dev/null 2>&1; then
  echo "build_ref: $OUT/emd*.so already built"
else
  SCRATCH="$(mktemp -d /tmp/ref_emd.XXXXXX)"
  cp "$REF"/extensions/emd/{emd.cpp,emd_cuda.cu,setup.py} "$SCRATCH"/
  ( cd "$SCRATCH" && TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=4 python setup.py -q build_ext --build-lib "$OUT" 2>/dev/null )
  rm -rf "$SCRATCH"
fi
ls -la "$OUT"
