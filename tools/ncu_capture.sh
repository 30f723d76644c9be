#!/bin/bash
# ncu --set full captures of single tcgen05 GEMM launches inside a warm full-size minibatch (gpurun; one GPU).
#   tools/ncu_capture.sh <tag>      -> gpurun_out/ncu_<tag>_fwd.ncu-rep (actor layer 2 forward, M=32768 N=K=1024)
#                                      gpurun_out/ncu_<tag>_dx.ncu-rep  (its dX GEMM: masked, column-summed, planes-only output)
#                                      gpurun_out/ncu_<tag>_dw.ncu-rep  (a split-K dW GEMM)
tag=${1:-r01}
common="--set full --import-source on --clock-control none --kernel-name-base demangled"
ncu $common --kernel-name "regex:gemm_tc256_kernel<.bool.0, .bool.0, .bool.1>" --launch-skip 28 --launch-count 1 -f -o gpurun_out/ncu_${tag}_fwd python tools/profile_minibatch.py 3 2 > gpurun_out/ncu_${tag}_fwd.log 2>&1
ncu $common --kernel-name "regex:gemm_tc256_kernel<.bool.0, .bool.1, .bool.1>" --launch-skip 30 --launch-count 1 -f -o gpurun_out/ncu_${tag}_dx python tools/profile_minibatch.py 3 2 > gpurun_out/ncu_${tag}_dx.log 2>&1
ncu $common --kernel-name "regex:gemm_tc256_kernel<.bool.1, .bool.1, .bool.1>" --launch-skip 32 --launch-count 1 -f -o gpurun_out/ncu_${tag}_dw python tools/profile_minibatch.py 3 2 > gpurun_out/ncu_${tag}_dw.log 2>&1
ls -la gpurun_out/*.ncu-rep
