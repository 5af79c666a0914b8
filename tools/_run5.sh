cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c_bench_8gpu.json 2> gpurun_out/r2c_bench_8gpu.err; echo "rc=$?"
tail -3 gpurun_out/r2c_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_8gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'])
for k in ('train_big4m','train'):
    t=d.get(k,{}); print(k, {x:t.get(x) for x in ('value','ms_per_step','error','exchange_bytes_in_per_rank_per_step','mean_loss','owner_sharded','replica_sync_ms_per_epoch','value_incl_epoch_sync')})
PY
for o in 1 0; do CHK_OWNER=$o timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 tools/dp_phase_times.py 2>&1 | grep -v "^W\|^\[W\|^$\|\*\*\*\|OMP_NUM" | tail -14 | tee -a gpurun_out/r2c_dp_phase_8gpu.log; done
