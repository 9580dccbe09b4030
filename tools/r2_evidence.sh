#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): tests, bench lines, march / adjoint at 1024^2,
# Euler / NS steps, assembly table, launch list and one full ncu capture of the top kernel.
# Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand afterwards.
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/r2_pytest_final.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_r2_2048.json 2> $O/bench_r2_2048.err
timeout 200 python bench.py --grid 1024 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r2_1024.json 2> $O/bench_r2_1024.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 --ref-budget 60 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
timeout 500 python tools/march.py --grid 1024 --steps 5 --adjoint > $O/r2_march_1024.txt 2>&1
timeout 200 python bench.py --workload euler --grid 256 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r2_euler_256.json 2> $O/bench_r2_euler_256.err
timeout 200 python bench.py --workload ns --grid 256 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r2_ns_256.json 2> $O/bench_r2_ns_256.err
timeout 200 python tools/kec_step.py --ni 384 --nj 384 --reps 1 > $O/r2_kec_euler_384.txt 2>&1
timeout 200 python tools/assembly_roofline.py --grid 2048 > $O/r2_assembly_roofline_2048.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
# launch list of one Newton step (only after the plain runs above have exited)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $O/r2_launches_2048.csv python tools/profile_step.py --grid 2048 > $O/r2_profile_step_ncu.log 2>&1
python tools/summarize_launches.py $O/r2_launches_2048.csv > $O/r2_launches_cavity2048_newton_step.txt 2>&1
gzip -f $O/r2_launches_2048.csv
# one full capture of the dominant kernel (fp32-operator instance of the TMA SpMV: the preconditioner)
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:k_spmv_tma -c 6 -o $O/r2_ncu_full_spmv python tools/profile_step.py --grid 2048 --only solve > $O/r2_ncu_full.log 2>&1
ncu -i $O/r2_ncu_full_spmv.ncu-rep --page details > $O/r2_ncu_full_k_spmv_tma_details.txt 2>&1
ncu -i $O/r2_ncu_full_spmv.ncu-rep --page raw --csv > $O/r2_ncu_full_k_spmv_tma_raw.csv 2>&1
ls -la $O | tail -30
tail -3 $O/r2_pytest_final.log
