#!/bin/bash
# multi-GPU evidence: N = $1 ranks -- the multi-rank pytest (torchrun outputs kept) and one bench line
N=$1
O=gpurun_out
mkdir -p $O
NPB_TEST_LOG_DIR=$O timeout 400 python -m pytest tests/test_gpu_multirank.py -q 2>&1 | tail -4 > $O/r2_pytest_multirank_${N}gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_r2_2048_N$N.out 2> $O/bench_r2_2048_N$N.err
grep "^{" $O/bench_r2_2048_N$N.out > $O/bench_r2_2048_N$N.json
cat $O/r2_pytest_multirank_${N}gpu.log | tail -2
python -c "
import json; d=json.load(open('$O/bench_r2_2048_N$N.json')); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['phases']['krylov_ms'], d['phases']['krylov_iters'])"
