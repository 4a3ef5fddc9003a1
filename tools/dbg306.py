import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from oim_b200 import build, lib, traces, abi
from oracle import bindings
import util
build.build(); bindings.build()
torch.zeros(1, device='cuda:0'); lib.init([0])
seed=306; nb=32768
t = traces.fuzz_trace(500, nb, seed=seed, max_io_blocks=[8, 64, 300, 1024][seed % 4], arena_bytes=(8 << 20))
want = util.run_oracle(bindings.PortOracle, t, nb)
for attempt in range(3):
    got = util.run_cuda(lib, t, nb, mem="host")
    bad = np.nonzero(got[2] != want[2])[0]
    print("attempt", attempt, "store diffs", len(bad), "arena diffs", int((got[1]!=want[1]).sum()))
    if len(bad):
        lo, hi = bad[0]//512, bad[-1]//512
        print("blocks", lo, hi, "first byte", bad[0], "got", got[2][bad[:4]], "want", want[2][bad[:4]])
        for i,r in enumerate(t.reqs):
            c=r['cdb']; op=c[0]
            if op in (0x28,0x2a): lba=int.from_bytes(bytes(c[2:6]),'big'); n=int.from_bytes(bytes(c[7:9]),'big')
            elif op in (0x88,0x8a): lba=int.from_bytes(bytes(c[2:10]),'big'); n=int.from_bytes(bytes(c[10:14]),'big')
            elif op in (0xa8,0xaa): lba=int.from_bytes(bytes(c[2:6]),'big'); n=int.from_bytes(bytes(c[6:10]),'big')
            elif op in (0x08,0x0a): lba=(int(c[1])<<16|int(c[2])<<8|int(c[3])); n=int(c[4]) or 256
            elif op==0x42: print(i, "UNMAP", "status", got[0][i]['status']); continue
            else: continue
            s,cn=int(r['iov_start']),int(r['iovcnt']); tot=int(t.iovs['len'][s:s+cn].sum())
            nn=max(n, tot//512)
            if lba<=hi and lba+nn>lo: print(i, i//32, hex(op), lba, n, "payload", tot, "iovcnt", cn, "dir", r['dir'], "status", got[0][i]['status'])
        break
lib.fini()
