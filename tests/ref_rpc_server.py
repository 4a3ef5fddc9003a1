"""Run the REFERENCE's JSON-RPC server (oracle/_ref/liboim_ref.so: S/lib/rpc + S/lib/jsonrpc + the
registered bdev / vhost handlers) as a process of its own:  python ref_rpc_server.py <rpc.sock> <vhost-dir> [vhost] [busy]
With the third argument the build that links the reference's own vhost-user transport is loaded
(liboim_ref_vhost.so): controllers then listen on <vhost-dir>/<name> for a vhost-user master."""
import ctypes as C
import os
import sys
import time

here = os.path.dirname(os.path.abspath(__file__))
name = "liboim_ref_vhost.so" if len(sys.argv) > 3 and sys.argv[3] == "vhost" else "liboim_ref.so"
ref = C.CDLL(os.path.join(os.path.dirname(here), "oracle", "_ref", name))
rc = ref.oimref_rpc_start(sys.argv[1].encode(), sys.argv[2].encode())
if rc != 0:
    sys.exit(f"oimref_rpc_start rc={rc}")
busy = "busy" in sys.argv[3:]        # a reactor never sleeps; the tests do not need a core burnt
while True:
    ref.oimref_rpc_poll(256 if busy else 64)
    if not busy:
        time.sleep(0.0005)
