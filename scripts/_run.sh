out=gpurun_out; tag=t51
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
python scripts/profile_solve.py 1024 50 2 3 2>&1 | grep "^rep" > $out/${tag}_base.log
for m in 5 6; do YR_LIB_OVERRIDE=$PWD/rootfinding_b200/csrc/_var/lib_m$m.so python scripts/profile_solve.py 1024 50 2 3 2>&1 | grep "^rep" > $out/${tag}_m$m.log; done
python scripts/profile_solve.py 16 50 2 3 2>&1 | grep "^rep" > $out/${tag}_b16.log
python scripts/profile_solve.py 1 50 2 3 2>&1 | grep "^rep" > $out/${tag}_b1.log
echo done
