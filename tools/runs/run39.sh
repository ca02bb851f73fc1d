for ph in 1 2; do
  R=$((ph*256+8))
  B2S_BA_REPEAT=$R timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_local_ba -s 1 -c 1 -f -o gpurun_out/ba_v16_rep$ph python tools/ba_prof.py 32 2 > gpurun_out/ncu_ba_rep$ph.log 2>&1
  tail -1 gpurun_out/ncu_ba_rep$ph.log | cut -c1-120
done
