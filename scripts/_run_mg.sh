out=gpurun_out; tag=${1:-mg}; N=${2:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $out/${tag}_scale$N.json 2> $out/${tag}_scale$N.err
echo "rc $?" >> $out/${tag}_scale$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > $out/${tag}_ref$N.json 2>> $out/${tag}_scale$N.err
echo done
