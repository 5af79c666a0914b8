cd /root/repo
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2b_dp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_dp_tests.log
tail -4 gpurun_out/r2b_dp_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r2b_dp_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_dp_check.log
grep -v "^W\|^\[W" gpurun_out/r2b_dp_check.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2b_bench_2gpu.json 2> gpurun_out/r2b_bench_2gpu.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b_bench_2gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'])
for k in ('train_big4m','train'):
    t=d.get(k,{}); print(k, {x:t.get(x) for x in ('value','ms_per_step','error','exchange_bytes_in_per_rank_per_step','mean_loss')})
PY
