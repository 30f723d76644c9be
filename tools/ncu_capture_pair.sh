#!/bin/bash
# ncu --set full captures of single persistent CTA-pair GEMM launches inside a warm full-size minibatch (gpurun; one GPU).
#   tools/ncu_capture_pair.sh <tag>  -> gpurun_out/ncu_<tag>_fwd.ncu-rep (forward: K-major x K-major), _dx (K-major x MN-major, bit mask + column sums), _dw (split-K)
tag=${1:-r02}
common="--set full --import-source on --clock-control none --kernel-name-base demangled"
ncu $common --kernel-name "regex:gemm_tc2_kernel<.bool.0, .bool.0>" --launch-skip 28 --launch-count 1 -f -o gpurun_out/ncu_${tag}_fwd python tools/profile_minibatch.py 3 2 > gpurun_out/ncu_${tag}_fwd.log 2>&1
ncu $common --kernel-name "regex:gemm_tc2_kernel<.bool.0, .bool.1>" --launch-skip 30 --launch-count 1 -f -o gpurun_out/ncu_${tag}_dx python tools/profile_minibatch.py 3 2 > gpurun_out/ncu_${tag}_dx.log 2>&1
ncu $common --kernel-name "regex:gemm_tc2_kernel<.bool.1, .bool.1>" --launch-skip 32 --launch-count 1 -f -o gpurun_out/ncu_${tag}_dw python tools/profile_minibatch.py 3 2 > gpurun_out/ncu_${tag}_dw.log 2>&1
ls -la gpurun_out/*${tag}*.ncu-rep
