#!/bin/bash
# One gpurun call that refreshes the round's evidence: GPU tests, bench line, ncu launch list (+ DRAM bytes per
# launch) and a full ncu capture of mid-solve tensor-kernel launches.  Outputs under gpurun_out/<tag>_*.
tag=${1:-ev}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
echo "tests rc $?" >> $out/${tag}_tests.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo "bench rc $?" >> $out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
python scripts/profile_solve.py 1024 50 2 2 > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1000 \
    --csv --log-file $out/${tag}_launches.csv python scripts/profile_solve.py 1024 50 2 2 > $out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:f2_tensor_warp -s 100 -c 8 \
    -o $out/${tag}_warp -f python scripts/profile_solve.py 1024 50 2 1 > $out/${tag}_ncu_full.log 2>&1
echo done
