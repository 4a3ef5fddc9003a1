set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-vu --no-lat --no-e2e --no-numa"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.out 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_queue_kernel -s 2 -c 1 -f -o gpurun_out/r2_rand4k $B --no-seq --no-vq --no-mixed --no-extra --no-sweep > gpurun_out/p1.out 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_queue_kernel -s 4 -c 1 -f -o gpurun_out/r2_vq $B --no-seq --no-mixed --no-extra --no-sweep > gpurun_out/p2.out 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_queue_kernel -s 4 -c 1 -f -o gpurun_out/r2_randwrite $B --no-seq --no-vq --no-mixed --no-sweep > gpurun_out/p3.out 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_shared_queue_kernel -s 26 -c 1 -f -o gpurun_out/r2_shared64 $B --no-seq --no-vq --no-mixed --no-extra > gpurun_out/p4.out 2>&1
ls -la gpurun_out/*.ncu-rep
