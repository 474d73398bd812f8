#!/usr/bin/env bash
# Test infrastructure: builds the UNMODIFIED reference kernels (anyprec.cu) for sm_100a into
# oracle/_ref/libapgemv_ref.so, straight from /root/reference (no sources are copied).
# Only runs where /root/reference exists (the build container); the .so travels to the GPU box.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF_ROOT:-/root/reference}/inference/ap_gemv"
OUT="$HERE/_ref"
[ -f "$REF/anyprec.cu" ] || { echo "reference not present at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"
if [ -f "$OUT/libapgemv_ref.so" ] && [ "$OUT/libapgemv_ref.so" -nt "$REF/anyprec.cu" ] && [ "$OUT/libapgemv_ref.so" -nt "$HERE/ref_shim.cu" ]; then
  echo "oracle/_ref/libapgemv_ref.so up to date"; exit 0
fi
TORCH_INC="$(python - <<'PY'
import os, torch
print(os.path.join(os.path.dirname(torch.__file__), "include"))
PY
)"
# anyprec.cu pulls ATen headers through datatype.h/typetraits.h but references no ATen symbol,
# so it links without libtorch.  Flags mirror the reference's setup.py:12-27 except the arch line.
nvcc -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC \
  -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ \
  -U__CUDA_NO_HALF2_OPERATORS__ -U__CUDA_NO_HALF2_CONVERSIONS__ \
  -U__CUDA_NO_BFLOAT16_OPERATORS__ -U__CUDA_NO_BFLOAT16_CONVERSIONS__ \
  -U__CUDA_NO_BFLOAT162_OPERATORS__ -U__CUDA_NO_BFLOAT162_CONVERSIONS__ \
  -gencode arch=compute_100a,code=sm_100a \
  -I"$REF" -I"$TORCH_INC" -I"$TORCH_INC/torch/csrc/api/include" \
  "$REF/anyprec.cu" "$HERE/ref_shim.cu" -o "$OUT/libapgemv_ref.so" -lcudart
echo "built $OUT/libapgemv_ref.so"
