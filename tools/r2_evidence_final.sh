#!/bin/bash
# final single-GPU evidence of round 2 (after the last solver changes)
O=gpurun_out
mkdir -p $O
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/r2_pytest_final.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_r2_2048.json 2> $O/bench_r2_2048.err
timeout 200 python bench.py --grid 1024 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r2_1024.json 2> $O/bench_r2_1024.err
timeout 400 python tools/march.py --grid 1024 --steps 5 --adjoint > $O/r2_march_1024.txt 2>&1
timeout 200 python bench.py --workload euler --grid 256 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r2_euler_256.json 2> $O/bench_r2_euler_256.err
timeout 200 python bench.py --workload ns --grid 256 --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_r2_ns_256.json 2> $O/bench_r2_ns_256.err
timeout 200 python tools/kec_step.py --ni 384 --nj 384 --reps 1 > $O/r2_kec_euler_384.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
tail -3 $O/r2_pytest_final.log; tail -3 $O/r2_march_1024.txt; cat $O/r2_smoke.log | tail -1
for f in bench_r2_2048 bench_r2_1024 bench_r2_euler_256 bench_r2_ns_256; do grep "^{" $O/$f.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['phases'].get('krylov_ms'), d['phases'].get('precond_setup_ms'))"; done
