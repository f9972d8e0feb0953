#!/bin/bash
bash tools/gpu_ka2.sh "$@"
MLX_PV_KA2=1 bash tools/gpu_ncu_pv.sh ka2c
