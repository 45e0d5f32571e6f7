#!/bin/bash
# timing diagnostics of the bit-packed GEMM: the product build next to builds with one part of the kernel removed
# (-DSBR_GB_<part>=0: LOAD = bit words by TMA, STS = bf16 expansion into the A stage, FENCE = proxy fence, TMA = B
# operand loads, EPI = epilogue stores; their results are wrong, only their timings are read)
mkdir -p gpurun_out
C=sibrar---single-branch-recommender_b200/csrc
V=$C/variants
mkdir -p $V
(cd $C && make -j8 > /dev/null 2>&1)
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
for v in LOAD STS FENCE TMA EPI ALL; do
  D="-DSBR_GB_$v=0"; [ $v = ALL ] && D="-DSBR_GB_LOAD=0 -DSBR_GB_STS=0 -DSBR_GB_TMA=0 -DSBR_GB_EPI=0"
  (nvcc $F $D -c $C/gemm_sm100.cu -o $V/g_$v.o 2> /dev/null &&
   nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $V/lib_no$v.so $V/g_$v.o $(ls $C/build/*.o | grep -v gemm_sm100) -lcudart) &
done
wait
{
echo "== product build"; timeout 120 python scripts/profile_gemm_bits.py
for v in LOAD STS FENCE TMA EPI ALL; do
  echo "== without $v"; SBR_LIB_PATH=$PWD/$V/lib_no$v.so timeout 120 python scripts/profile_gemm_bits.py
done
} > gpurun_out/gemm_bits_variants.log 2>&1
cat gpurun_out/gemm_bits_variants.log
rm -rf $V
