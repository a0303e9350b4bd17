#!/bin/bash
# Multi-GPU evidence on N GPUs of one node (run under `gpurun --gpus N`): sharded-frontier correctness, strong scaling of one
# frontier, weak scaling of the bench.  Outputs gpurun_out/<tag>_*.  MODE=weak runs the weak-scaling bench only.
N=${1:-2}; MODE=${2:-all}; out=gpurun_out; tag=r2mg$N
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$MODE" = "all" ]; then
$TR --master-port 29611 scripts/sharded_check.py 2>&1 | grep "world=\|EQUAL\|MISMATCH" > $out/${tag}_check.log; tail -4 $out/${tag}_check.log | cut -c1-300
for st in round_robin rank0; do
$TR --master-port 29612 bench.py --mode strong --gpus $N --steps 3 --warmup 2 --strong-systems 1024 --strong-degree 50 --strong-start $st > $out/${tag}_strong_1024_$st.json 2> $out/${tag}_strong.err; cut -c1-700 $out/${tag}_strong_1024_$st.json
done
$TR --master-port 29613 bench.py --mode strong --gpus $N --steps 3 --warmup 2 --strong-systems 8 --strong-degree 100 --strong-start rank0 > $out/${tag}_strong_8x100.json 2>> $out/${tag}_strong.err; cut -c1-700 $out/${tag}_strong_8x100.json
fi
$TR --master-port 29614 bench.py --gpus $N --steps 10 --warmup 3 > $out/${tag}_weak.json 2> $out/${tag}_weak.err; cut -c1-900 $out/${tag}_weak.json
echo done
