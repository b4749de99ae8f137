#!/usr/bin/env python3
"""A packed Stwo record whose transcript has to REPEAT a felt draw (channel.simf:115-141: one of the first four words of the drawn digest is >= 2p;
probability 2^-29 per draw, so none of the reference's fixtures exercises it).

  python tests/golden/make_retry_fixture.py [--threads 8]     # ~2^29 trials of 3 compressions: a few minutes of host time with SHA-NI

Takes the golden witnesses (both presets), keeps everything but the trace root (COMMITMENTS.1), and searches trace roots {original words 0..5, counter} until
the cp_alpha draw (evals/commit.simf:29) fails its uniformity test at least once.  The resulting proofs are of course rejected (their trace root is not the
root of the decommitted tree) — what they pin is the retry bookkeeping: the counter n_sent, the number of draws, every later digest, draw and query of the
transcript, identical in the oracle and on the GPU (tests/test_draw_retry.py).  Writes tests/golden/draw_retry_root.json (the eight words of the trace root; the tests splice it into the packed golden records)."""
import argparse
import ctypes as C
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import witparse as W  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    args = ap.parse_args()
    orc = O.Oracle()
    fn = orc.lib.oracle_grind_draw_retry
    fn.restype = C.c_int
    found_root = None
    roots = {}
    for preset in ("testing", "prod"):
        p = O.PRESETS[preset]
        wit = W.load_wit(open(os.path.join(HERE, f"stwo_proof_{preset}.wit")).read())
        packed, rej = W.pack_stwo(wit, p["n_queries"], p["n_fri_layers"], p["lde_log"])
        assert not rej
        cfg = O.make_config(preset, O.MODE_PROVER_CONSISTENT)
        lo = orc.stwo_layout(cfg)
        # channel state before the trace root is mixed = digest after mixing COMMITMENTS.0 into the zero state (both fixtures share COMMITMENTS.0 = SHA-256(""))
        c0 = packed[lo.off_commit:lo.off_commit + 8]
        state = orc.channel_mix_u256(np.zeros(9, dtype=np.uint32), O.words_u256(c0))  # (digest[8], n_sent)
        dig = np.ascontiguousarray(state[:8], dtype=np.uint32)
        base = packed[lo.off_commit + 8:lo.off_commit + 16].copy()
        if found_root is None or not (packed[lo.off_commit:lo.off_commit + 8] == first_c0).all():
            first_c0 = c0.copy()
            chunk = 1 << 24

            def scan(k):
                out = (C.c_uint32 * 8)()
                ok = fn(dig.ctypes.data_as(O.u32p), base.ctypes.data_as(O.u32p), C.c_uint64(k * chunk), C.c_uint64(chunk), out)
                return np.array(out, dtype=np.uint32) if ok else None

            found_root, k0 = None, 0
            with ThreadPoolExecutor(args.threads) as ex:
                while found_root is None:
                    for r in ex.map(scan, range(k0, k0 + args.threads)):
                        if r is not None and found_root is None:
                            found_root = r
                    k0 += args.threads
                    print(f"{preset}: scanned {k0 * chunk:,} roots", flush=True)
        rec = packed.copy()
        rec[lo.off_commit + 8:lo.off_commit + 16] = found_root  # the whole root: both fixtures share COMMITMENTS.0, so one root serves both presets
        _, status, tr = orc.stwo_verify_batch(cfg, rec, 1, want_trace=True)
        assert tr[0].draw_retries >= 1, "the search found nothing"
        roots[preset] = [int(x) for x in found_root]
        print(f"{preset}: trace root words 6..7 = {found_root[6]:#x} {found_root[7]:#x}, draw_retries = {tr[0].draw_retries}, status = {status[0]:#x}")
    import json

    assert roots["testing"] == roots["prod"]
    json.dump({"trace_root_words": roots["prod"], "note": "COMMITMENTS.1 (the trace root, most significant word first) for which the cp_alpha draw of both golden "
               "witnesses is repeated once: replace words off_commit + 8 .. + 16 of the packed golden record"}, open(os.path.join(HERE, "draw_retry_root.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
