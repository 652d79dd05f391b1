#!/bin/bash
# Round 2, GPU call W: what the pieces of the host-facing call cost (PCIe floor microbenchmark), the call timed from C.
mkdir -p gpurun_out
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv > gpurun_out/r2w_pcie.txt 2>&1; cat gpurun_out/r2w_pcie.txt
nproc; lscpu | grep -i "model name\|numa" | head -5
scripts/micro/pcie_floor > gpurun_out/r2w_pcie_floor.txt 2>&1; cat gpurun_out/r2w_pcie_floor.txt
MTX=$(python -c "from sextans_b200 import workloads as w; print(w.suitesparse_path('nasa4704'))")
for g in 1 2 4; do scripts/micro/e2e_c $MTX 16 $g; done 2>&1 | tee gpurun_out/r2w_e2e_c.txt
scripts/micro/e2e_c $MTX 16 1 0 2>&1 | tee -a gpurun_out/r2w_e2e_c.txt
