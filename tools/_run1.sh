set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_train.py -x -q -k "pair_coef" > gpurun_out/r2b_coef_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_coef_tests.log
tail -5 gpurun_out/r2b_coef_tests.log
timeout 300 python tools/reduce_bench.py coef > gpurun_out/r2b_reduce_coef.log 2>&1
cat gpurun_out/r2b_reduce_coef.log
for wl in fb237 big4m wn18rr yago310; do
  timeout 200 python tools/train_profile.py $wl >> gpurun_out/r2b_tp.log 2>&1
  CHK_PAIR_COEF=0 timeout 200 python tools/train_profile.py $wl >> gpurun_out/r2b_tp.log 2>&1
done
cat gpurun_out/r2b_tp.log
