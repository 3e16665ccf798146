#!/bin/bash
# Timing-only A/B visit (no tests): interleaved runs of the listed builds on the bench workload.
mkdir -p gpurun_out
timeout 900 python tools/ab_libs.py 592 ${AB_REPS:-3} $AB_LIBS > gpurun_out/ab.txt 2>&1
cat gpurun_out/ab.txt
