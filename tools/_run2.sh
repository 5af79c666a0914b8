cd /root/repo
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_gpu_tests.log
tail -8 gpurun_out/r2b_gpu_tests.log
for wl in fb237 big4m; do
CHK_PROFILE_EAGER_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2b_train_${wl}_launches.csv python tools/train_profile.py $wl 4 > gpurun_out/r2b_ncu_${wl}.log 2>&1
done
python tools/launch_summary.py gpurun_out/r2b_train_fb237_launches.csv gpurun_out/r2b_train_big4m_launches.csv
