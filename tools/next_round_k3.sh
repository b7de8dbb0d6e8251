#!/usr/bin/env bash
# First GPU session of the next round: validate and time the experimental register-based rare path of K3
# (csrc/gemm_topk.cu, VS_RARE_REGS) against the committed kernel.  Run HERE (builds two libraries with nvcc,
# then one gpurun call for each; ~70 GPU-seconds per call):
#
#     bash tools/next_round_k3.sh
#
# Adopt VS_RARE_REGS=1 as the default only if tests/test_gemm_gpu.py passes with it AND the probe's `base`
# line beats the committed kernel's (2.10 ms K3 kernel time at 10 M x 128, batch 1024, profiles/r01_k3_probe_experiments.txt).
set -euo pipefail
cd "$(dirname "$0")/.."
run() {
  /usr/local/graft/bin/gpurun --timeout 200 -- \
    "timeout 40 python tools/k3_probe.py base: tauinf:B200VS_GEMM_TAU_INF=1 > gpurun_out/$1_probe.txt 2>&1; cat gpurun_out/$1_probe.txt; timeout 120 python -m pytest tests/test_gemm_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -3"
}
(cd mlx-vector-db_b200 && B200VS_RARE_REGS=1 python -m b200vs.build --force)
run k3_rare_regs
(cd mlx-vector-db_b200 && python -m b200vs.build --force)       # back to the committed default
run k3_default
