#!/bin/bash
# Turns the outputs of scripts/gpu_evidence.sh (gpurun_out/<tag>_*) into the tracked files profiles/<round>_*.
tag=${1:-ev}; rnd=${2:-r02}; out=gpurun_out; P=profiles
tail -1 $out/${tag}_bench.json > $P/${rnd}_bench.json
tail -1 $out/${tag}_bench_ref.json > $P/${rnd}_bench_reference.json
cp $out/${tag}_launches.csv $P/${rnd}_launches.csv
{
echo "ncu launch lists (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum; --clock-control none; serialised launches,"
echo "cold caches: the SHARES are what carries over to the un-profiled run).  scripts/launch_shares.py."
echo; echo "== two solves of the bench workload: 1024 random 2-D degree-50 systems (scripts/profile_solve.py 1024 50 2 2)"
python scripts/launch_shares.py $out/${tag}_launches.csv 2>/dev/null | head -16
echo; echo "== one solve of 512 random 3-D degree-10 systems (scripts/profile_solve.py 512 10 3 1)"
python scripts/launch_shares.py $out/${tag}_launches_3d.csv 2>/dev/null | head -8
echo; echo "== one solve of 512 random 4-D degree-4 systems (scripts/profile_solve.py 512 4 4 1)"
python scripts/launch_shares.py $out/${tag}_launches_4d.csv 2>/dev/null | head -7
} > $P/${rnd}_launch_shares.txt
python scripts/traffic_json.py $P/${rnd}_launches.csv $P/${rnd}_bench.json 2 "f2_tensor_(warp|cta)" > $P/${rnd}_traffic.json
{
echo "Device-resident solve times of the BASELINE.json configurations on one B200 (scripts/profile_solve.py B deg dim reps; last"
echo "repetition; boxes/s = boxes / wall time of run_jobs, inputs in HBM; 'GB/s algorithmic' = 16 B x coefficients of every box at"
echo "entry / summed tensor-kernel time).  2-D systems with extents <= 79 and all 3-D / 4-D systems run on the frontier pipelines"
echo "(DESIGN.md 2, 2.4); 2-D degree 100 starts on the level pipeline and hands over; 5-D runs on the level pipeline.  The memory"
echo "lines are the library pools' high-water mark after the solves (what ChebyshevSubdivisionSolver._PEAK_FACTOR is fitted to)."
echo
cut -c1-400 $out/${tag}_configs.log | sed 's/^rep [0-9]: //' | awk 'NR==FNR{print; next}' 
echo; echo "(the last line: use_fma = 1, 2-D degree 50: fused multiply-add in the contraction; roots within 1e-10, restrictions no longer bit-identical)"
} > $P/${rnd}_configs.txt
{
echo "Latency of yr.solve (host callables in, roots out) on Chebfun2 systems and the notebook demos, one B200 vs the CPU oracle"
echo "(scripts/solve_latency.py; second call of each; 'device solves' = frontier solves, 'regions' = boxes of the re-solve loop"
echo "of Combined_Solver.solve approximated and solved together)."
echo
cat $out/${tag}_latency.log | grep -v Warning | grep -v "warnings.warn" | cut -c1-220
} > $P/${rnd}_solve_latency.txt
tail -3 $out/${tag}_tests.log > $P/${rnd}_gpu_tests.txt
tail -2 $out/${tag}_smoke.log | cut -c1-400 >> $P/${rnd}_gpu_tests.txt
ls -la $P/${rnd}_*
