#!/bin/bash
# debug_build.sh : ad_mpc_b200/variants/debug.so = every kernel file compiled with -DADMPC_DEBUG (device-side bounds asserts,
# common.cuh ADMPC_ASSERT).  Run the parity suite under it:  ADMPC_LIB=ad_mpc_b200/variants/debug.so python -m pytest tests -m gpu
set -e
cd "$(dirname "$0")/.."
scripts/build_variant.sh debug "-DADMPC_DEBUG" $(cd ad_mpc_b200/csrc && ls *.cu)
