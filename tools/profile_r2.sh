# ncu captures kept under profiles/r2_* (run under gpurun on one B200; numbers printed by a run under ncu are never bench values)
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-vu --no-lat --no-e2e --no-numa --no-poller-leg"
# every launch of OUR kernels with its device time (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu -k regex:oim_ --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_launches.out 2>&1
# headline leg: 254 queues x 16512 requests, one CTA per queue (third launch)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_queue_kernel -s 2 -c 1 -f -o gpurun_out/r2_rand4k $B --no-seq --no-vq --no-mixed --no-extra --no-sweep > gpurun_out/p1.out 2>&1
# virtqueue leg: 4096 guest rings in HBM (oim_lun_vring_kernel, second launch)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_vring_kernel -s 1 -c 1 -f -o gpurun_out/r2_vq $B --no-seq --no-mixed --no-extra --no-sweep > gpurun_out/p2.out 2>&1
# 4 KiB random write, 254 queues (launch 5 of oim_lun_queue_kernel: headline x3, then the write leg)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_queue_kernel -s 4 -c 1 -f -o gpurun_out/r2_randwrite $B --no-seq --no-vq --no-mixed --no-sweep > gpurun_out/p3.out 2>&1
# queue sharing at 64 queues (sweep legs of 1,2,4,8,16 queues x 5 launches come first)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:oim_lun_shared_queue_kernel -s 26 -c 1 -f -o gpurun_out/r2_shared64 $B --no-seq --no-vq --no-mixed --no-extra > gpurun_out/p4.out 2>&1
ls -la gpurun_out/*.ncu-rep
