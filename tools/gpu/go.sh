#!/bin/bash
# build, then run a script on the GPU box: tools/gpu/go.sh [--gpus N] <timeout_s> <script> 
set -e
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
make -C /root/repo/rustradio_b200/csrc -j8 -s 2>&1 | grep -v "warning #550\|bool force\|\^$\|^$\|Remark" || true
make -C /root/repo/oracle -s
gpurun $G --timeout $1 -- bash $2
