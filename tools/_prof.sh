cd /root/repo
for wl in big4m fb237; do
CHK_PROFILE_EAGER_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"reduce_apply|score_gather" --launch-skip 6 -c 2 -o gpurun_out/r2_final_train_${wl} -f python tools/train_profile.py $wl 4 > gpurun_out/r2_final_ncu_${wl}.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
