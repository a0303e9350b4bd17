out=gpurun_out; tag=t55
for m in 0 1 2; do YR_F2_SCALAR=$m python scripts/profile_solve.py 1024 50 2 3 2>&1 | grep "^rep [12]" > $out/${tag}_s$m.log; done
YR_F2_SCALAR=1 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
echo done
