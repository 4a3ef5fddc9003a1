# which of today's changes cost the virtqueue legs 4-6 %?  (old = the tree before them)
BF="--steps 10 --warmup 3 --no-seq --no-e2e --no-cpu --no-mixed --no-lat --no-vu --no-sweep --no-extra"
for v in default old noust noustpad; do
  if [ $v = default ]; then unset OIM_LIB_PATH; else export OIM_LIB_PATH=$PWD/oim_b200/liboimgpu_$v.so; fi
  timeout 150 python bench.py $BF 2> gpurun_out/b5_$v.err | grep "^{" > gpurun_out/b5_$v.json
done
ls -la gpurun_out/b5_*.json
