#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --mode joint --kernel-table --no-roofline > gpurun_out/r2c_$tag.json 2> gpurun_out/r2c_$tag.err; echo "$tag $(cut -c1-118 gpurun_out/r2c_$tag.json)"; grep -E "^vs_(conv3x3x3_tc_kdn_ex|inorm_relu_bwd_reduce|conv3x3x3_dgrad) " gpurun_out/r2c_$tag.err; }
run default A=1
run nofuse VAESEG_NO_FUSE_REDUCE=1
run fuse48 VAESEG_FUSE_REDUCE_MAX_VOX=110592
run fuse24 VAESEG_FUSE_REDUCE_MAX_VOX=13824
run default2 A=1
