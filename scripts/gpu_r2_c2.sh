#!/bin/bash
set -u
bash scripts/gpu_r2_dbg.sh
bash scripts/gpu_r2_c.sh
