out=gpurun_out; tag=t59
python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1
python scripts/sanitize_case.py > $out/${tag}_plain.log 2>&1 &&
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitize_case.py > $out/${tag}_memcheck.log 2>&1
echo "memcheck rc $?" >> $out/${tag}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python scripts/sanitize_case.py > $out/${tag}_racecheck.log 2>&1
echo "racecheck rc $?" >> $out/${tag}_racecheck.log
python scripts/profile_e2e.py > $out/${tag}_e2e.log 2>&1
python bench.py --no-cpu-baseline > $out/${tag}_bench.json 2> $out/${tag}_bench.err
echo done
