#!/bin/bash
# Runs tests marked `gpu` on a machine WITHOUT a GPU, through the unchanged C ABI over a simulated CUDA runtime
# (racon_b200/csrc/cuda_sim_runtime.h).  Results only — it is slow (a 500-base window takes seconds) and
# says nothing about speed.  Pick small tests:
#   tools/sim_gpu_tests.sh tests/test_zz_gpu_resident.py
#   tools/sim_gpu_tests.sh tests/test_zz_files_pipeline.py -k fastq_sam
#   tools/sim_gpu_tests.sh tests/test_gpu_poa.py -k "edge_cases or batch_object_protocol"
set -e
cd "$(dirname "$0")/.."
python -c "from racon_b200 import build; print(build.build_simapi())"
export RACON_B200_LIB="$PWD/racon_b200/lib/simapi/libracon_b200.so"
exec python -m pytest -m gpu -x -q "$@"
