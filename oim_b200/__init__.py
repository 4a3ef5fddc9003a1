"""oim_b200 — B200-native block-I/O data path behind intel/oim's SPDK plugin surface.

  csrc/        hand-written sm_100a kernels + the C ABI (liboimgpu.so, include/oimgpu.h)
  daemon/      oim-gpu-vhost: the JSON-RPC daemon the unmodified OIM Go binaries talk to
  lib.py       ctypes binding (mirrors pkg/spdk's shim names)
  abi.py       wire structures and CDB builders
  traces.py    bdevperf-shaped and adversarial request traces
  vring.py     virtio split rings in a guest-memory image
  build.py     in-tree nvcc / g++ builds
"""
