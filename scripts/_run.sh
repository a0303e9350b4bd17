out=gpurun_out; tag=t52
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
python scripts/profile_e2e.py > $out/${tag}_e2e.log 2>&1
python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo done
