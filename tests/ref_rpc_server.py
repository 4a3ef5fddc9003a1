"""Run the REFERENCE's JSON-RPC server (oracle/_ref/liboim_ref.so: S/lib/rpc + S/lib/jsonrpc + the
registered bdev / vhost handlers) as a process of its own:  python ref_rpc_server.py <rpc.sock> <vhost-dir>"""
import ctypes as C
import os
import sys
import time

here = os.path.dirname(os.path.abspath(__file__))
ref = C.CDLL(os.path.join(os.path.dirname(here), "oracle", "_ref", "liboim_ref.so"))
rc = ref.oimref_rpc_start(sys.argv[1].encode(), sys.argv[2].encode())
if rc != 0:
    sys.exit(f"oimref_rpc_start rc={rc}")
while True:
    ref.oimref_rpc_poll(64)
    time.sleep(0.0005)
