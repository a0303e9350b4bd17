out=gpurun_out; tag=r2f
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 scripts/sharded_check.py 2>&1 | grep "world=\|EQUAL\|MISMATCH" > $out/${tag}_check2.log; tail -3 $out/${tag}_check2.log
python bench.py --mode strong --gpus 1 --steps 3 --warmup 2 --strong-systems 1024 --strong-degree 50 > $out/${tag}_strong1_1024.json 2> $out/${tag}_strong1.err; cut -c1-330 $out/${tag}_strong1_1024.json
for st in round_robin rank0; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --mode strong --gpus 2 --steps 3 --warmup 2 --strong-systems 1024 --strong-degree 50 --strong-start $st > $out/${tag}_strong2_1024_$st.json 2> $out/${tag}_strong2.err; cat $out/${tag}_strong2_1024_$st.json
done
echo done
