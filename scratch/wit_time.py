import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stark_symphony_b200 as S
cfg = S.stwo_config("prod", 1)
raw = open("tests/golden/stwo_proof_prod.wit","rb").read()
ver = S.Verifier(0)
for n in (148, 512, 2048):
    blob = np.tile(np.frombuffer(raw, dtype=np.uint8), n)
    offs = np.arange(n+1, dtype=np.uint64)*np.uint64(len(raw))
    d_blob = torch.from_numpy(blob).cuda(); d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    for rep in range(3):
        torch.cuda.synchronize(); t0=time.perf_counter()
        p, f = ver.stwo_pack_wit_batch(d_blob, d_offs, cfg)
        torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print(f"device pack n={n}: {dt*1e3:.3f} ms  -> {n/dt:.0f} wit/s, {n*len(raw)/dt/1e9:.1f} GB/s text", flush=True)
    assert int(f.sum().item())==0
