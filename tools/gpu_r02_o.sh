#!/bin/bash
set -u
mkdir -p gpurun_out
DMV_FUSE_FC_ADAM=0 timeout 600 python tools/timeline.py gpurun_out/r02o_timeline_unfused.txt > gpurun_out/r02o_tl0.log 2>&1; echo "tl0 exit $?"
DMV_FUSE_FC_ADAM=1 DMV_FC_ADAM_VARIANT=9 timeout 600 python tools/timeline.py gpurun_out/r02o_timeline_fused9.txt > gpurun_out/r02o_tl9.log 2>&1; echo "tl9 exit $?"
tail -5 gpurun_out/r02o_tl0.log
