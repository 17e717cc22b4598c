#!/bin/bash
bash tools/gpu_multi.sh 2 2>&1 | tail -12
