#!/bin/bash
bash tools/gpu_profile_pred_int8.sh
bash tools/gpu_profile_int8.sh
