#!/bin/bash
# timing diagnostics of the bit-packed GEMM: the product build next to builds with one part of the kernel removed
# (csrc/variants/lib_no*.so, built with -DSBR_GB_<part>=0; their results are wrong, only their timings are read)
mkdir -p gpurun_out
V=sibrar---single-branch-recommender_b200/csrc/variants
{
echo "== product build"; timeout 120 python scripts/profile_gemm_bits.py
for v in LOAD STS FENCE TMA EPI ALL; do
  echo "== without $v"; SBR_LIB_PATH=$PWD/$V/lib_no$v.so timeout 120 python scripts/profile_gemm_bits.py
done
} > gpurun_out/r02_gemm_bits_variants.log 2>&1
cat gpurun_out/r02_gemm_bits_variants.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 3 -c 1 -f \
  -o gpurun_out/r02_gemm_bits python scripts/profile_gemm_bits.py > gpurun_out/r02_gemm_bits_ncu.log 2>&1
tail -2 gpurun_out/r02_gemm_bits_ncu.log
