"""One scan-shaped workload for an `ncu --set full` capture: a synthetic store of the given shape and a few
64-query top-10 batches.  Usage: python scripts/ncu_scan_shape.py ROWS DTYPE [DIM] [NQ]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import vidmem_b200 as vm
from vidmem_b200.store import EmbeddingStore

rows, dt = int(sys.argv[1]), sys.argv[2]
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 384
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 64
st = EmbeddingStore(dim, rows, dt)
st.synth_fill(3, rows)
st.set_size(rows)
g = torch.Generator().manual_seed(7)
q = (torch.randint(-127, 128, (nq, dim), generator=g).float() / 128.0).cuda()
for _ in range(4):
    st.topk_device(q, 10, flags=vm.VM_FLAG_ASYNC)
torch.cuda.synchronize()
print("ok", rows, dt, dict(st.counters()))
st.close()
