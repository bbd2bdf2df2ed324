#!/bin/bash
# Staged GPU validation; every stage is its own process with its own timeout so one trap cannot hide the rest.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout "${TMO:-600}" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n "${TAILN:-15}" gpurun_out/$name.log; }
run kern_nogemm python -m pytest tests/test_kernels_gpu.py -q -x -k "not gemm"
run mstcn python -m pytest tests/test_mstcn_gpu.py -q -x -s
TAILN=60 run first_contact python scripts/first_contact.py
run kern_gemm python -m pytest tests/test_kernels_gpu.py -q -k "gemm"
TAILN=40 run evp python -m pytest tests/test_evp_gpu.py -q -s
run smoke python __graft_entry__.py smoke
