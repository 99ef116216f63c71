#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/pageable_e2e.py 2>&1 | tail -8 | tee gpurun_out/r2_pageable_e2e.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
