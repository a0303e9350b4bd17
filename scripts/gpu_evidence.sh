#!/bin/bash
# One gpurun call that refreshes the round's evidence: GPU tests, both bench arms, the configurations table, the solve
# latencies, ncu launch lists (+ DRAM bytes per launch) of the 2-D bench workload and of a 3-D / 4-D frontier solve.
# Outputs under gpurun_out/<tag>_*; scripts/collect_evidence.sh turns them into profiles/<round>_*.
tag=${1:-ev}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
echo "tests rc $?" >> $out/${tag}_tests.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc $?" >> $out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1
{
for cfg in "4096 20 2 3" "1024 50 2 3" "512 75 2 2" "256 100 2 3" "512 10 3 3" "512 4 4 3" "64 6 4 2" "256 2 5 2"; do
  python scripts/profile_solve.py $cfg 2>&1 | grep "^rep\|device memory" | tail -2
done
YR_FMA=1 python scripts/profile_solve.py 1024 50 2 3 2>&1 | grep "^rep" | tail -1
} > $out/${tag}_configs.log 2>&1
python scripts/solve_latency.py > $out/${tag}_latency.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1100 \
    --csv --log-file $out/${tag}_launches.csv python scripts/profile_solve.py 1024 50 2 2 > $out/${tag}_ncu_list.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 \
    --csv --log-file $out/${tag}_launches_3d.csv python scripts/profile_solve.py 512 10 3 1 > $out/${tag}_ncu_list3.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 \
    --csv --log-file $out/${tag}_launches_4d.csv python scripts/profile_solve.py 512 4 4 1 > $out/${tag}_ncu_list4.log 2>&1
echo done
