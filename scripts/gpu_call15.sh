#!/bin/bash
O=gpurun_out/r02; mkdir -p $O
timeout 120 ./scripts/ubench/umma_rate > $O/umma_rate.log 2>&1; echo "umma rc=$?"; cat $O/umma_rate.log
