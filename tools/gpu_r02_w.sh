#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"thin_conv|thin_deconv|thin_wgrad_mma" -c 6 -f -o gpurun_out/thin_prof python tools/time_layers.py "5s2 224" > gpurun_out/r02w_ncu_thin.log 2>&1; echo "ncu thin exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fc_wgrad_adam" -c 2 -f -o gpurun_out/fcadam_prof env DMV_FC_ADAM_VARIANT=9 python tools/time_fc_adam.py > gpurun_out/r02w_ncu_fcadam.log 2>&1; echo "ncu fcadam exit $?"
