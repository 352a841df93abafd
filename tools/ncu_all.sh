#!/bin/bash
# One `ncu --set full` capture per kernel family (run on the GPU box AFTER the plain commands exited 0).
# Reports land in gpurun_out/r01_full_<name>.ncu-rep; summarise them here with tools/ncu_summary.py.
set -u
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
cap() {  # name, kernel regex, launch-skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  $N -k "regex:$k" -s "$skip" -c 1 -o gpurun_out/r01_full_$name "$@" > gpurun_out/r01_full_$name.log 2>&1
  echo "$name rc=$?"
}
cap stem            '^stem_kernel'            2  python tools/profile_step.py 3
cap pair128         'conv_igemm2_kernel'      0  python tools/dom_conv.py 4 20 128 128
cap tcn             'tcn_igemm_kernel'        24 python tools/profile_step.py 3
cap fusion          'fusion_head_kernel'      2  python tools/profile_step.py 3
cap vgg_stem        'vgg_stem_pool_kernel'    2  python tools/vggish_bench.py 2400 1
cap maxpool         'maxpool2x2_kernel'       8  python tools/vggish_bench.py 2400 1
cap preproc         'preprocess_kernel'       2  python tools/preproc_bench.py 2400 1
cap row_gemm        'row_gemm_kernel'         200 python tools/train_bench.py 16 1
cap wgrad           'wgrad_kernel'            200 python tools/train_bench.py 16 1
cap logmel          'logmel_kernel'           0  python tools/logmel_bench.py
