set -u
cd lpopc_b200/csrc
for rows in 1 2 4; do for c in 1 4 5 6; do
  sed -i "s/static constexpr int SWEEP_ROWS = [0-9]*; static constexpr int SWEEP_MIN_CTAS = [0-9]*;/static constexpr int SWEEP_ROWS = $rows; static constexpr int SWEEP_MIN_CTAS = $c;/" ../../include/problems/synthetic20.h
  make lpb_problem_LpbSynthetic20.o > /dev/null 2>&1 && make > /dev/null 2>&1
  echo "rows=$rows ctas=$c $(grep -A3 Lb1ELb1ELb0ELb1E lpb_problem_LpbSynthetic20.ptxas.log | grep -o 'Used [0-9]* registers')"
  (cd ../..; python scripts/kernel_sweep.py synthetic20 1 --intervals 10000 --nodes 10 --unroll -1 2>&1 | tail -1)
done; done
