N=${1:-2}; out=gpurun_out; tag=r2n$N
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29611 scripts/sharded_check.py 2>&1 | grep "world=\|EQUAL\|MISMATCH\|rror" > $out/${tag}_check.log; tail -14 $out/${tag}_check.log | cut -c1-330
$TR --master-port 29612 bench.py --mode strong --gpus $N --steps 3 --warmup 2 --strong-systems 1024 --strong-degree 50 --strong-start round_robin > $out/${tag}_strong_1024.json 2> $out/${tag}_strong.err; cut -c1-420 $out/${tag}_strong_1024.json; tail -3 $out/${tag}_strong.err
$TR --master-port 29613 bench.py --mode strong --gpus $N --steps 3 --warmup 2 --strong-systems 8 --strong-degree 100 --strong-start rank0 > $out/${tag}_strong_8x100.json 2>> $out/${tag}_strong.err; cut -c1-420 $out/${tag}_strong_8x100.json
