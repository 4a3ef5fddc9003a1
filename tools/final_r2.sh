# the round's closing GPU call: whole GPU suite and the default bench line (tools/profile_r2_sg.sh and the memcheck
# pass of profiles/r2_sanitizer.md ran on the same queue kernels one call earlier)
timeout 500 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -8 > gpurun_out/t5.log
timeout 400 python bench.py --steps 10 --warmup 3 2> gpurun_out/b5.err | grep "^{" > gpurun_out/r2_bench_n1_final.json
tail -3 gpurun_out/t5.log
