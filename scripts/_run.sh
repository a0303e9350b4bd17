out=gpurun_out; tag=t57
YR_PROFILE=2.5 python scripts/solve_latency.py > $out/${tag}_lat.log 2>&1
echo done
