out=gpurun_out; tag=t61
python scripts/profile_solve.py 1024 50 2 3 2>&1 | grep "^rep [12]" > $out/${tag}_base.log
export YR_LIB_OVERRIDE=$PWD/rootfinding_b200/csrc/_var/lib_pair.so
python scripts/profile_solve.py 1024 50 2 3 2>&1 | grep "^rep [12]" > $out/${tag}_pair.log
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "oracle or golden or full_size or chebfun" > $out/${tag}_tests_pair.log 2>&1
echo done
