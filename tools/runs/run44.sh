for k in k_blur k_resize k_quadtree k_bow_resolve; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_v16_$k python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-150
done
