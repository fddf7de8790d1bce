#!/bin/bash
# tuning sweep of the bulk-copy SpMM (rebuilds spmm.cu on the GPU box per configuration)
for cfg in "4 2 4096 4" "4 2 2048 4" "4 2 2048 2" "4 2 1024 2" "4 3 2048 2" "8 2 2048 2" "4 2 4096 2" "4 2 2048 1"; do
  set -- $cfg
  touch c2dsr_b200/csrc/spmm.cu
  C2DSR_NVCC_DEFS="-DC2DSR_SPMM_WARPS=$1 -DC2DSR_SPMM_STAGES=$2 -DC2DSR_SPMM_STAGE_BYTES=$3 -DC2DSR_SPMM_RPW=$4" python -m c2dsr_b200.build > /dev/null 2>gpurun_out/build_err.txt || { echo "build failed $cfg"; tail -3 gpurun_out/build_err.txt; continue; }
  echo "== warps=$1 stages=$2 stage_bytes=$3 rpw=$4"
  timeout 200 python scratch/spmm_bench.py 2>&1 | grep "fwd\|bwd"
done
