#!/bin/bash
# final single-GPU evidence of round 2 (after the last solver changes): tests, bench lines, march +
# adjoint at 1024^2, Euler / NS steps, smoke, launch list (warm caches) and one full ncu capture
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_pytest_final.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_r2_2048.json 2> $O/bench_r2_2048.err
timeout 200 python bench.py --grid 1024 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r2_1024.json 2> $O/bench_r2_1024.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 --ref-budget 60 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
timeout 400 python tools/march.py --grid 1024 --steps 5 --adjoint > $O/r2_march_1024.txt 2>&1
timeout 200 python bench.py --workload euler --grid 256 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r2_euler_256.json 2> $O/bench_r2_euler_256.err
timeout 200 python bench.py --workload ns --grid 256 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r2_ns_256.json 2> $O/bench_r2_ns_256.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
timeout 100 python tools/gs_bench.py > $O/r2_gs_bench.txt 2>&1
timeout 100 python tools/dia_bench.py > $O/r2_dia_bench.txt 2>&1
# launch list of one Newton step (only after the plain runs above have exited)
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r2_launches_2048.csv python tools/profile_step.py --grid 2048 > $O/r2_profile_step_ncu.log 2>&1
python tools/summarize_launches.py $O/r2_launches_2048.csv > $O/r2_launches_cavity2048_newton_step.txt 2>&1
gzip -f $O/r2_launches_2048.csv
# one full capture of the dominant kernels (the Jacobian SpMV and the fp32 instances of the preconditioner)
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:k_spmv_tma -c 6 -o $O/r2_ncu_full_spmv python tools/profile_step.py --grid 2048 --only solve > $O/r2_ncu_full.log 2>&1
ncu -i $O/r2_ncu_full_spmv.ncu-rep --page details > $O/r2_ncu_full_k_spmv_tma_details.txt 2>&1
ncu -i $O/r2_ncu_full_spmv.ncu-rep --page raw --csv > $O/r2_ncu_full_k_spmv_tma_raw.csv 2>&1
tail -3 $O/r2_pytest_final.log; tail -3 $O/r2_march_1024.txt; cat $O/r2_smoke.log | tail -1
for f in bench_r2_2048 bench_r2_1024 bench_r2_euler_256 bench_r2_ns_256; do grep "^{" $O/$f.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['phases'].get('krylov_ms'), d['phases'].get('precond_setup_ms'))"; done
