#!/bin/bash
for cfg in "8 1024" "1 0" "2 1024"; do
  set -- $cfg
  echo "== TSR_CLUSTER=$1 TSR_STAGE_BUDGET=$2"
  TSR_CLUSTER=$1 TSR_STAGE_BUDGET=$2 timeout 400 python scripts/profile_phases.py c3_shot1532_41755_expansions c3_surface_d11_2048 c2_color_d7_shot877_1.2M_pushes c2_color_d7_first600 bb144_p1e-3_longbeam_256 | cut -c1-430
done
