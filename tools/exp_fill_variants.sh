BF="--steps 10 --warmup 3 --no-e2e --no-cpu --no-vq --no-mixed --no-lat --no-vu --no-poller-leg --no-sweep"
OIM_LIB_PATH=$PWD/oim_b200/liboimgpu_stream.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_shared_queues.py tests/test_vring.py tests/test_poller.py tests/test_multi_target.py -m gpu -q --timeout 150 2>&1 | tail -8 > gpurun_out/t3_stream.log
for v in default stream f128 f512 f128s f64s; do
  if [ $v = default ]; then unset OIM_LIB_PATH; else export OIM_LIB_PATH=$PWD/oim_b200/liboimgpu_$v.so; fi
  timeout 120 python bench.py $BF 2> gpurun_out/b3_$v.err | grep "^{" > gpurun_out/b3_$v.json
done
tail -3 gpurun_out/t3_stream.log
