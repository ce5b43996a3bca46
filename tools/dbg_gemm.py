import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ctc_asr_b200 import ops, _lib
C = _lib.COMPUTE_BF16X3
def run(M,N,K,ta,tb,seed):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    b = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    c = ops.gemm(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), ta=ta, tb=tb, compute=C).cpu().numpy().astype(np.float64)
    want = (a.T if ta else a).astype(np.float64) @ (b.T if tb else b).astype(np.float64)
    err = np.abs(c-want); i = np.unravel_index(err.argmax(), err.shape)
    print(M,N,K,ta,tb, "max rel err %.3e at %s; rows with err>1e-4*max: %s" % (err.max()/np.abs(want).max(), i, np.unique(np.where(err > 1e-4*np.abs(want).max())[0])[:20]), "cols:", np.unique(np.where(err > 1e-4*np.abs(want).max())[1])[:20])
run(1000,264,1048,False,False,1265)
run(1000,264,1048,False,False,1265)
run(384,512,256,False,False,1)
run(1000,264,1048,False,False,1265)
run(1000,264,1056,False,False,3)
run(1000,256,1048,False,False,3)
run(1000,264,1024,False,False,3)
