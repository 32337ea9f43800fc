#!/bin/bash
# One GPU call that validates and times the three opt-in paths prepared at the end of round 1:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/check_experimental.sh 2>&1 | tail -80'
# Each tool prints parity against the default path / the oracle first and timings second; every step has its own timeout so
# that a hang in one kernel cannot take the GPU box with it.
cd "$(dirname "$0")/.."
echo "=== DSU_RIC_PIXEL_MAJOR (conv_ric_persist_kernel<4>) ==="; timeout 150 python tools/ric_px_check.py 2>&1 | tail -12
echo "=== DSU_RIC_FIRST (conv_ric_first.cu) ===";                timeout 150 python tools/ric_first_check.py 2>&1 | tail -10
echo "=== DSU_SUBPIXEL (conv_halo_persist_kernel<true>) ===";     timeout 200 python tools/subpixel_check.py 2>&1 | tail -16
