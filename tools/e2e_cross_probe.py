#!/usr/bin/env python3
"""End-to-end rate of version 3 compact records verified under ANOTHER mode than they were packed under (include/ssym.h: two passes over the
kernels per chunk), next to the same-mode legs.  One GPU, pinned host buffers, enqueue-only calls (ssym_set_host_async), copies inside the clock.

  python tools/e2e_cross_probe.py [passes=64] [steps=6]
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import stark_symphony_b200 as S


def main():
    passes = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    n = 1024
    ver = S.Verifier(0)
    cfgs = {"ref-literal": S.stwo_config("prod", S.MODE_REF_LITERAL), "prover-consistent": S.stwo_config("prod", S.MODE_PROVER_CONSISTENT)}
    text = open(os.path.join(ROOT, "tests", "golden", "stwo_proof_prod.wit")).read()
    one, bad = S.witness.pack_stwo_wits([text], cfgs["ref-literal"])
    assert not bad[0]
    host_batch = np.tile(one, n)
    words = (n + 31) // 32
    bound = int(S.load().ssym_stwo_compact_bound(C.byref(cfgs["ref-literal"]), n))

    def pack(cfg):
        buf = torch.empty(bound, dtype=torch.int32).pin_memory()
        full, off = S.witness.compact_stwo(host_batch, cfg, out=buf.numpy().view(np.uint32), ver=ver)
        off_t = torch.empty(n + 1, dtype=torch.int64).pin_memory()
        off_t.numpy()[:] = off.view(np.int64)
        return buf, full[: int(off[n])], off_t.numpy().view(np.uint64), int(off[n]) * 4 // n

    blobs = {k: pack(c) for k, c in cfgs.items()}
    rows = torch.zeros((passes, words), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    out = {}
    for packed_under in cfgs:
        for verify_under, cfg in cfgs.items():
            _, blob, off, bytes_per_proof = blobs[packed_under]
            call = lambda acc: ver.stwo_verify_compact_batch(blob, off, cfg, accept_out=acc)
            ref = np.zeros(words, dtype=np.uint32)
            for _ in range(3):
                call(ref)
            t0 = time.perf_counter()
            for _ in range(32):
                call(ref)  # blocking calls
            sync_rate = n * 32 / (time.perf_counter() - t0)
            ver.set_host_async(True)
            for k in range(4):
                call(rows[k])
            ver.synchronize()
            rows[:] = 0xA5A5A5A5
            t0 = time.perf_counter()
            enq = 0.0
            for _ in range(steps):
                h0 = time.perf_counter()
                for p in range(passes):
                    call(rows[p])
                enq += time.perf_counter() - h0
                ver.synchronize()
            secs = time.perf_counter() - t0
            ver.set_host_async(False)
            assert (rows == ref[None, :]).all()
            accepted = int(np.unpackbits(ref.view(np.uint8), bitorder="little")[:n].sum())
            out[f"packed {packed_under} / verified {verify_under}"] = {
                "proofs_per_s": n * passes * steps / secs, "blocking_calls_proofs_per_s": sync_rate, "bytes_per_proof": bytes_per_proof, "h2d_gbs": n * passes * steps * bytes_per_proof / secs / 1e9,
                "host_enqueue_share": enq / secs, "accepted": accepted}
    print(json.dumps(out, indent=1))
    ver.close()


if __name__ == "__main__":
    main()
