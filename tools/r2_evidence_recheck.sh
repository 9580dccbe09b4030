O=gpurun_out
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_pytest_final.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_r2_2048.json 2> $O/bench_r2_2048.err
timeout 200 python bench.py --grid 1024 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_r2_1024.json 2> $O/bench_r2_1024.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1
timeout 100 python tools/profile_setup.py --grid 2048 > $O/r2_setup_profile_2048.txt 2>&1
tail -2 $O/r2_pytest_final.log; tail -1 $O/r2_smoke.log; tail -12 $O/r2_setup_profile_2048.txt
for f in bench_r2_2048 bench_r2_1024; do grep "^{" $O/$f.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['phases'].get('krylov_ms'), d['phases'].get('precond_setup_ms'), d['phases'].get('assembly_ms'))"; done
