set -x
for w in k1 k2g huge push; do
  timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitizer_run.py $w 2>&1 | grep -v "^$" | tail -3
done
for w in k1 k2g; do
  timeout 280 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitizer_run.py $w 2>&1 | grep -v "^$" | tail -3
done
