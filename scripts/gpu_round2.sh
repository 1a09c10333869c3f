#!/bin/bash
# GPU-box script of round 2: parity tests, bench (both arms), ncu launch list, full captures of the three node kernels,
# per-configuration latency table.  Outputs under gpurun_out/; scripts/summarise_profile.py turns them into profiles/r02_*.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu --no-extras"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_cons_jac -s 3 -c 1 -f -o gpurun_out/prof_cons_jac $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu cons_jac rc=$?"
CMD3="python scripts/dev/hess_sweep.py quadrotor 4096 --variants 0 --steps 2"
$CMD3 > gpurun_out/hs.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hess_tiled -s 2 -c 1 -f -o gpurun_out/prof_hess_tiled $CMD3 > gpurun_out/ncu_hess.log 2>&1
echo "ncu hess rc=$?"
CMD4="python scripts/kernel_sweep.py synthetic20 1 --intervals 10000 --nodes 10 --steps 3"
$CMD4 > gpurun_out/ks.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_cons_jac_rows -s 2 -c 1 -f -o gpurun_out/prof_rows $CMD4 > gpurun_out/ncu_rows.log 2>&1
echo "ncu rows rc=$?"
CMD5="python scripts/dev/kkt_probe.py 1024"
$CMD5 > gpurun_out/kkt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_kkt_factor -s 1 -c 1 -f -o gpurun_out/prof_kkt $CMD5 > gpurun_out/ncu_kkt.log 2>&1
echo "ncu kkt rc=$?"
python scripts/dev/solver_kernels.py 4096 > gpurun_out/solver_kernels.txt 2>&1; echo "solver kernels rc=$?"
python scripts/dev/permesh_probe.py > gpurun_out/permesh.txt 2>&1; echo "permesh rc=$?"
python scripts/config_table.py > gpurun_out/config_table.txt 2> gpurun_out/config_table.err; echo "config table rc=$?"
python scripts/parity_report.py > gpurun_out/parity_report.txt 2> gpurun_out/parity_report.err; echo "parity report rc=$?"
ls -la gpurun_out | tail -20
