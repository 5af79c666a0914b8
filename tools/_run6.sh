cd /root/repo
N=${1:-2}
timeout 400 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/r2d_dp_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_dp_tests.log; tail -3 gpurun_out/r2d_dp_tests.log
for px in 1; do for g in 0 1; do
CHK_PEER_EXCHANGE=$px CHK_GRAPH=$g timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tools/dp_phase_times.py 2>&1 | grep -v "^W\|^\[W\|^$\|\*\*\*\|OMP_NUM\|NCCL version" | tail -14 | tee -a gpurun_out/r2d_dp_phase_${N}gpu.log
done; done
