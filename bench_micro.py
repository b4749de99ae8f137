#!/usr/bin/env python3
"""Secondary measurements (BASELINE.json configs 1, 4, 5-shape) on ONE B200.

  python bench_micro.py [--quick] [--out profiles/r01_micro.json]

* config 4: element-wise jets over 2^28 uniformly random canonical elements (HBM roofline) and the Merkle-path sweep
  (2^20 paths, depth 16..28; INT32 roofline; 32d+68 bytes and 2d compressions per path);
* config 1: the stark101 proof replicated x65536 (+ corrupted negatives), device resident;
* large-batch Stwo (2^14, 2^16 proofs) to show where the 1024-proof headline sits relative to the asymptote.
Every timing: CUDA events on the launching stream, >= 3 warm-up launches, best of 5.  Results are spot-checked against
the oracle (test infrastructure) outside the timed regions.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

P = 2147483647


def timed(fn, stream, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="2^24 elements / 2^16 paths (smoke)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import stark_symphony_b200 as S
    from oracle import oracle as O

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    ver = S.Verifier(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ver.set_stream(stream.cuda_stream)
    orc = O.Oracle()
    int32_peak, _ = ver.int32_peak_probe()
    results = {"hbm_peak_gbs": hbm_peak, "int32_peak_lanes_per_s": int32_peak, "field": [], "merkle_sweep": [], "stark101": None, "stwo_large": []}
    g = torch.Generator(device="cuda").manual_seed(0)  # Philox, seed 0

    # ---- config 4a: field jets over N elements --------------------------------------------------------------
    N = 1 << (24 if args.quick else 28)

    def rnd(words):
        return torch.randint(0, P, (N * words,), dtype=torch.int32, device="cuda", generator=g)

    a1, b1 = rnd(1), rnd(1)
    a4, b4 = rnd(4), rnd(4)
    pos = (torch.randint(0, 1 << 20, (N,), dtype=torch.int32, device="cuda", generator=g) & ~1)
    cases = [
        ("m31_mul", lambda: ver.m31_mul(a1, b1), 12),
        ("m31_inv", lambda: ver.m31_inv(a1), 8 + 1),
        ("qm31_mul", lambda: ver.qm31_mul(a4, b4), 48),
        ("qm31_inv", lambda: ver.qm31_inv(a4), 32 + 1),
        ("circle_fold(log 20, table twiddles)", lambda: ver.circle_fold(pos, a4, b4, a4, 20), 4 + 16 * 3 + 16 + 1),
        ("line_fold(log 20, table twiddles)", lambda: ver.line_fold(pos, a4, b4, a4, 20), 4 + 16 * 3 + 16 + 1),
    ]
    for name, fn, bytes_per in cases:
        ms = timed(fn, stream)
        gbs = N * bytes_per / (ms * 1e-3) / 1e9
        results["field"].append({"op": name, "elements": N, "ms": ms, "elements_per_s": N / (ms * 1e-3), "bytes_per_element": bytes_per,
                                 "gb_per_s": gbs, "frac_of_hbm_peak": gbs / hbm_peak})
        print(f"{name:40s} {ms:9.3f} ms  {N / ms / 1e6:9.2f} G elem/s  {gbs:8.1f} GB/s  ({gbs / hbm_peak * 100:5.1f} % of measured HBM peak)", flush=True)
    # spot check against the oracle
    out = ver.qm31_mul(a4, b4).cpu().numpy().view(np.uint32)
    ha, hb = a4[:64].cpu().numpy().view(np.uint32), b4[:64].cpu().numpy().view(np.uint32)
    for i in range(16):
        assert list(out[4 * i:4 * i + 4]) == list(orc.qm31_mul(ha[4 * i:4 * i + 4], hb[4 * i:4 * i + 4]))
    o2, _ = ver.circle_fold(pos, a4, b4, a4, 20)
    o2 = o2.cpu().numpy().view(np.uint32)
    hp = pos[:16].cpu().numpy().view(np.uint32)
    for i in range(8):
        w, _ = orc.circle_fold(int(hp[i]), ha[4 * i:4 * i + 4], hb[4 * i:4 * i + 4], 20, ha[4 * i:4 * i + 4])
        assert list(o2[4 * i:4 * i + 4]) == list(w)
    del a1, b1, a4, b4, pos, out, o2
    torch.cuda.empty_cache()

    # ---- config 4b: Merkle path sweep -----------------------------------------------------------------------
    M = 1 << (16 if args.quick else 20)
    leaf = torch.randint(-2**31, 2**31 - 1, (M * 8,), dtype=torch.int32, device="cuda", generator=g)
    for depth in (16, 20, 24, 28):
        sib = torch.randint(-2**31, 2**31 - 1, (M * depth * 8,), dtype=torch.int32, device="cuda", generator=g)
        auth = (torch.randint(0, 1 << depth, (M,), dtype=torch.int64, device="cuda", generator=g) + (1 << depth)).to(torch.int32)
        roots, _, _ = ver.merkle_root_from_path(leaf, auth, sib, depth)  # tree-free construction: the computed roots become the expected ones
        ms = timed(lambda: ver.merkle_root_from_path(leaf, auth, sib, depth, expected_root=roots), stream)
        _, _, ok = ver.merkle_root_from_path(leaf, auth, sib, depth, expected_root=roots)
        assert int((ok.cpu().numpy().view(np.uint32) != 0xFFFFFFFF).sum()) == 0  # every path verifies
        # one path against the oracle
        hl, hs, ha_ = leaf[:8].cpu().numpy().view(np.uint32), sib[:depth * 8].cpu().numpy().view(np.uint32), int(auth[0].item()) & 0xFFFFFFFF
        _, croot, fpath = orc.merkle_verify_32(O.words_u256(hl), ha_, [O.words_u256(hs[8 * k:8 * k + 8]) for k in range(depth)], 0)
        assert fpath == 1 and list(O.u256_words(croot)) == list(roots[:8].cpu().numpy().view(np.uint32))
        bytes_per = 32 * depth + 68
        pair_hashes = M * depth / (ms * 1e-3)
        alu_lanes = pair_hashes * 1677  # SHF + LOP3 instructions of one pair hash (adds are IMADs on the FMA pipe)
        results["merkle_sweep"].append({"depth": depth, "paths": M, "ms": ms, "paths_per_s": M / (ms * 1e-3), "pair_hashes_per_s": pair_hashes,
                                        "compressions_per_s": 2 * pair_hashes, "gb_per_s": M * bytes_per / (ms * 1e-3) / 1e9,
                                        "frac_of_hbm_peak": M * bytes_per / (ms * 1e-3) / 1e9 / hbm_peak,
                                        "alu_pipe_frac": alu_lanes / int32_peak,
                                        "literal_int_ops_per_s": 2 * pair_hashes * 2296, "literal_frac_of_int32_peak": 2 * pair_hashes * 2296 / int32_peak})
        print(f"merkle depth {depth}: {ms:8.3f} ms  {M / ms / 1e3:8.2f} M paths/s  {pair_hashes / 1e9:6.2f} G pair-hashes/s  "
              f"{M * bytes_per / ms / 1e6:7.1f} GB/s  ALU pipe {alu_lanes / int32_peak * 100:5.1f} %", flush=True)
        del sib, auth, roots
        torch.cuda.empty_cache()
    del leaf

    # ---- config 1: stark101 ----------------------------------------------------------------------------------
    golden = os.path.join(ROOT, "tests", "golden")
    blob, offs, bad = S.witness.pack_stark101_wits([open(os.path.join(golden, "stark101_proof.wit")).read()])
    n = 1 << (12 if args.quick else 16)
    all_blob = np.tile(blob, n)
    bad_rows = list(range(7, n, 97))
    for r in bad_rows:
        all_blob[r * len(blob) + 5] += 1  # wrong last layer
    offsets = np.arange(n + 1, dtype=np.uint64) * len(blob)
    d_blob = torch.from_numpy(all_blob.view(np.int32)).cuda()
    d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
    ms = timed(lambda: ver.stark101_verify_batch(d_blob, d_off), stream)
    accept, _, _ = ver.stark101_verify_batch(d_blob, d_off)
    ver.synchronize()
    bits = np.unpackbits(accept.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
    expect = np.ones(n, dtype=bool)
    expect[bad_rows] = False
    assert (bits == expect).all()
    t0 = time.perf_counter()
    orc.s101_verify_batch(all_blob[: 64 * len(blob)], offsets[:65])
    cpu_per = (time.perf_counter() - t0) / 64
    results["stark101"] = {"proofs": n, "ms": ms, "proofs_per_s": n / (ms * 1e-3), "compressions_per_s": n * 480 / (ms * 1e-3),
                           "packed_bytes_per_proof": int(len(blob) * 4), "gb_per_s": n * len(blob) * 4 / (ms * 1e-3) / 1e9,
                           "oracle_1_thread_proofs_per_s": 1 / cpu_per}
    print(f"stark101 x{n}: {ms:8.3f} ms  {n / ms / 1e3:8.2f} M proofs/s  (oracle, 1 thread: {1 / cpu_per:8.0f} proofs/s)", flush=True)
    del d_blob, d_off
    # the same batch as `.wit` TEXT (the file `make run` hands to simfony), tokenised on the GPU: device resident and from pinned host memory
    raw = open(os.path.join(golden, "stark101_proof.wit"), "rb").read()
    t_pinned = torch.empty(len(raw) * n, dtype=torch.uint8).pin_memory()
    t_np = t_pinned.numpy()
    t_np.reshape(n, len(raw))[:] = np.frombuffer(raw, dtype=np.uint8)
    t_off = np.arange(n + 1, dtype=np.uint64) * np.uint64(len(raw))
    d_text, d_toff = t_pinned.cuda(), torch.from_numpy(t_off.view(np.int64)).cuda()
    best_dev, best_host = 1e30, 1e30
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acc_t, _, _ = ver.stark101_verify_wit_batch(d_text, d_toff)
        torch.cuda.synchronize()
        if rep:
            best_dev = min(best_dev, time.perf_counter() - t0)
        t0 = time.perf_counter()
        acc_h, _, _ = ver.stark101_verify_wit_batch(t_np, t_off)
        if rep:
            best_host = min(best_host, time.perf_counter() - t0)
    assert (np.unpackbits(acc_h.view(np.uint8), bitorder="little")[:n] == 1).all() and bool((acc_t == -1).all().item())
    results["stark101"]["from_wit_text"] = {"text_bytes_per_proof": len(raw), "device_resident_proofs_per_s": n / best_dev, "host_pinned_proofs_per_s": n / best_host,
                                            "host_text_gb_per_s": n * len(raw) / best_host / 1e9,
                                            "note": "ssym_stark101_verify_wit_batch, synchronous call, wall clock (one H2D of the whole text, no chunk overlap)"}
    print(f"stark101 x{n} from .wit text: {n / best_dev / 1e6:.2f} M proofs/s device resident, {n / best_host / 1e6:.2f} M proofs/s from pinned host text", flush=True)
    del d_text, d_toff, t_pinned

    # ---- Stwo at larger batches -------------------------------------------------------------------------------
    cfg = S.stwo_config("prod", S.MODE_REF_LITERAL)
    lo = S.stwo_layout(cfg)
    packed, _ = S.witness.pack_stwo_wits([open(os.path.join(golden, "stwo_proof_prod.wit")).read()], cfg)
    one = torch.from_numpy(packed.view(np.int32)).cuda()
    for logn in ((10, 12) if args.quick else (10, 12, 14, 16)):
        n = 1 << logn
        dev = one.repeat(n)
        acc = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
        ver.set_pipeline_depth(1)
        ms = timed(lambda: ver.stwo_verify_batch(dev, cfg, n, accept_out=acc), stream)
        results["stwo_large"].append({"proofs": n, "ms": ms, "proofs_per_s": n / (ms * 1e-3), "compressions_per_s": n * 3806 / (ms * 1e-3),
                                      "gb_per_s": n * 54488 / (ms * 1e-3) / 1e9, "note": "serial launches (pipeline depth 1), device resident"})
        print(f"stwo x{n}: {ms:9.3f} ms  {n / ms / 1e3:8.3f} M proofs/s (serial)", flush=True)
        del dev
        torch.cuda.empty_cache()
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)
    ver.close()


if __name__ == "__main__":
    main()
