#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --mode joint --no-roofline > gpurun_out/r2i_$tag.json 2>/dev/null; echo "$tag $(cut -c1-118 gpurun_out/r2i_$tag.json)"; }
run off A=1
run fuse12 VAESEG_FUSE_APPLY=1 VAESEG_FUSE_APPLY_MAX_VOX=1728
run fuse24 VAESEG_FUSE_APPLY=1 VAESEG_FUSE_APPLY_MAX_VOX=13824
run fuse6 VAESEG_FUSE_APPLY=1 VAESEG_FUSE_APPLY_MAX_VOX=216
run off2 A=1
run fuse12b VAESEG_FUSE_APPLY=1 VAESEG_FUSE_APPLY_MAX_VOX=1728
