#!/bin/bash
# ncu --set full of the staged kernel on the C5-like power-law matrix: the unwindowed launch
# and the four column-window passes (first / middle / middle / last)
mkdir -p gpurun_out
PROBE_W=0,262144 PROBE_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on \
  -k regex:spmm_staged -s 2 -c 5 -o gpurun_out/r1_win_prof python scripts/probe_windows.py > gpurun_out/r1_win_prof.log 2>&1
echo "ncu exit $?"; tail -5 gpurun_out/r1_win_prof.log; ls -la gpurun_out/
