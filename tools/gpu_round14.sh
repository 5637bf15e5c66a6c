#!/bin/bash
# Source-level ncu capture of the heavy training kernels only (a whole step is > 64 MiB of report).
TAG=${1:-r2v}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
sha256sum videopose3d_b200/_lib/libvp3d_b200.so | cut -d' ' -f1 > gpurun_out/${TAG}_lib_sha256.txt
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
  --kernel-name-base demangled \
  -k 'regex:conv_gemm_kernel<\(int\)256, \(bool\)1|wgrad_gemm_kernel<\(int\)256|bn_bwd_apply|adam_pack' \
  -o gpurun_out/${TAG}_full_train_bf16 -f python tools/profile_steps.py train bf16 > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log
ls -la gpurun_out/${TAG}_full_train_bf16.ncu-rep
