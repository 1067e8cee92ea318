#!/bin/bash
# ab.sh B N M variant lib... : QP kernel time of each library variant (ADMPC_LIB) on one config, for A/B runs in one gpurun call
B=$1; N=$2; M=$3; V=$4; shift 4
for lib in "$@"; do
  echo "== $lib"
  ADMPC_LIB=$lib ADMPC_QP_VARIANT=$V python - <<PY
import sys, os
sys.path.insert(0, os.getcwd())
sys.argv = ["sweep"]
from scripts.sweep import run
run($B, $N, $M, $V)
PY
done
