#!/bin/bash
# One GPU call = one fixed overhead (box acquisition + snapshot push, 1-3 GPU-minutes charged), so batch the work:
#   gpurun --timeout 1500 -- 'bash tools/gpu_cycle.sh [quick|full|profile]'
# quick   : parity tests of both kernels + one timing line each              (~1.5 min of run time)
# full    : the whole GPU suite, smoke, bench.py (full line), aligner timings (~5 min)
# profile : quick + ncu captures of both kernels and the bench launch list   (~8 min)
# Everything lands in gpurun_out/ (merged back by gpurun); numbers printed under ncu are never bench values.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}" || exit 1
mode="${1:-quick}"
mkdir -p gpurun_out
run() { echo "== $*"; timeout "${T:-600}" "$@" 2>&1 | tail -"${N:-3}"; }
if [ "$mode" = full ]; then
  T=1200 run python -m pytest tests -x -q -m gpu
  T=120 run python -c "import __graft_entry__ as g; g.smoke()"
  timeout 400 python bench.py 2>gpurun_out/bench.err | tail -1 > gpurun_out/bench.json; tail -c 400 gpurun_out/bench.json; echo
else
  T=900 run python -m pytest tests/test_gpu_poa.py tests/test_gpu_aln.py tests/test_breaking_points.py tests/test_pipeline.py -x -q -m gpu
  N=1 run python tools/profile_poa.py --windows 10000 --launches 2
fi
N=1 run python tools/bench_aln.py --pairs 12000 --len 8000
N=1 run python tools/bench_aln.py --pairs 60000 --len 1000
if [ "$mode" = profile ]; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:rp_poa_kernel -s 1 -c 1 -f -o gpurun_out/poa \
      python tools/profile_poa.py --windows 4736 > gpurun_out/ncu_poa.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:rp_aln_kernel -c 1 -f -o gpurun_out/aln \
      python tools/bench_aln.py --pairs 6000 --len 8000 --reps 1 > gpurun_out/ncu_aln.log 2>&1
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
fi
