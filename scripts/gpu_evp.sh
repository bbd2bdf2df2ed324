#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; echo "=== $name" ; timeout "${TMO:-900}" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n "${TAILN:-15}" gpurun_out/$name.log; }
TAILN=60 run evp python -m pytest tests/test_evp_gpu.py -q -s -x
run smoke python __graft_entry__.py smoke
