# ncu captures of the byte-granular SG legs (bench.py seq128k_sg), kept under profiles/r2_unaligned*_ncu.md
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu --no-vu --no-lat --no-e2e --no-numa --no-poller-leg --no-vq --no-mixed --no-sweep"
# launches of oim_lun_queue_kernel: headline 0-2, seq pages 3-5, single 6-8, unaligned 9-11, unaligned+3 12-14, scattered 15-17
timeout 400 ncu --set full --clock-control none --import-source on -k regex:oim_lun_queue_kernel -s 13 -c 1 -f -o gpurun_out/r2_unaligned3 $B > gpurun_out/p5.out 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:oim_lun_queue_kernel -s 16 -c 1 -f -o gpurun_out/r2_scattered $B > gpurun_out/p6.out 2>&1
ls -la gpurun_out/*.ncu-rep
