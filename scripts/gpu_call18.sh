#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 120 ./scripts/ubench/epi_rate > $O/epi_rate.log 2>&1; echo "rc=$?"; cat $O/epi_rate.log
