"""A transcript that has to REPEAT a felt draw (channel.simf:115-141: one of the first four words of the drawn digest is >= 2p — once in 2^29 draws, so no
fixture of the reference exercises it).  tests/golden/draw_retry_root.json holds a trace root, found by search (tests/golden/make_retry_fixture.py), for
which the cp_alpha draw of both golden witnesses is repeated once.  The records are rejected (the root is not the decommitted tree's); what they pin is the
retry bookkeeping — the counter n_sent, every later digest, draw and query — identical in the oracle, in the cost model and on the GPU."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O
from test_oracle_fixtures import load_stwo

ROOT_WORDS = np.array(json.load(open(os.path.join(GOLDEN, "draw_retry_root.json")))["trace_root_words"], dtype=np.uint32)


def retry_record(orc, preset):
    rec = load_stwo(preset).copy()
    lo = orc.stwo_layout(O.make_config(preset, 0))
    rec[lo.off_commit + 8:lo.off_commit + 16] = ROOT_WORDS
    return rec, lo


def be(words):
    return np.asarray(words, dtype=">u4").tobytes()


@pytest.mark.parametrize("preset", ["testing", "prod"])
def test_oracle_repeats_the_draw_as_channel_simf_says(orc, preset):
    rec, lo = retry_record(orc, preset)
    _, status, tr = orc.stwo_verify_batch(O.make_config(preset, O.MODE_PROVER_CONSISTENT), rec, 1, want_trace=True)
    t = tr[0]
    assert t.draw_retries == 1 and status[0] != 0 and not (status[0] & 1)  # repeated, not exhausted
    # independent restatement with hashlib: digest after mixing the two roots; first draw (counter 0) is not uniform, the second (counter 1) is taken
    d = hashlib.sha256(bytes(32) + be(rec[lo.off_commit:lo.off_commit + 8])).digest()
    d = hashlib.sha256(d + be(ROOT_WORDS)).digest()
    first = np.frombuffer(hashlib.sha256(d + (0).to_bytes(4, "big")).digest(), dtype=">u4")
    second = np.frombuffer(hashlib.sha256(d + (1).to_bytes(4, "big")).digest(), dtype=">u4")
    assert (first[:4] >= 4294967294).any() and (second[:4] < 4294967294).all()
    assert list(t.cp_alpha) == [int(x) % 2147483647 for x in second[:4]]
    # the composition root is mixed into the SAME digest (a draw does not move it) and the counter restarts
    d2 = hashlib.sha256(d + be(rec[lo.off_commit + 16:lo.off_commit + 24])).digest()
    assert bytes(be(t.digest_commit)) == d2


def test_cost_model_counts_the_repeated_draw(orc):
    import stark_symphony_b200 as S
    from test_cost_model import model_counts, oracle_counts

    for preset in ("testing", "prod"):
        rec, _ = retry_record(orc, preset)
        cfg = O.make_config(preset, O.MODE_REF_LITERAL)
        want, tr, _ = oracle_counts(orc, cfg, rec)
        assert tr.draw_retries == 1
        assert model_counts(S, cfg, list(tr.queries)[: cfg.n_queries], tr.draw_retries) == want
        base, _, _ = oracle_counts(orc, cfg, load_stwo(preset))
        assert want[0] == base[0] + 1 and want[-1] == 1  # one more compression than the golden witness, one retry


@pytest.mark.gpu
@pytest.mark.parametrize("preset", ["testing", "prod"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_gpu_transcript_repeats_the_draw(orc, preset, mode):
    """The warp-specialised transcript kernel prepares the next message before it has seen the current draw; a repeated draw is its one mis-prediction.
    A batch that mixes records with and without a repeated draw (the 32 transcripts of a CTA run in lockstep: the others hash along) gives the oracle's
    trace for every one of them."""
    import stark_symphony_b200 as S

    rec, lo = retry_record(orc, preset)
    golden = load_stwo(preset)
    other = golden.copy()
    other[lo.off_pow_nonce + 1] += 1
    cfg = S.stwo_config(preset, mode & 1, dedup_queries=bool(mode & 2))
    ocfg = O.make_config(preset, mode)
    for pattern in ([rec], [golden, rec, other], [rec] * 33 + [golden] * 31 + [rec, other, rec], [golden] * 40 + [rec]):
        batch = np.concatenate(pattern)
        ver = S.Verifier(0)
        accept, status, traces = ver.stwo_verify_batch(batch, cfg, len(pattern), want_status=True, want_trace=True)
        # and without traces, pipelined: the status words alone
        import torch

        dev = torch.from_numpy(batch.view(np.int32)).cuda()
        ver.set_pipeline_depth(4)
        sts = [ver.stwo_verify_batch(dev, cfg, len(pattern), want_status=True)[1] for _ in range(5)]
        ver.synchronize()
        ver.close()
        _, o_status, o_traces = orc.stwo_verify_batch(ocfg, batch, len(pattern), want_trace=True)
        assert (status == o_status).all()
        assert all((s_.cpu().numpy().view(np.uint32) == o_status).all() for s_ in sts)
        for i in range(len(pattern)):
            assert bytes(memoryview(traces[i]).cast("B")) == bytes(memoryview(o_traces[i]).cast("B")), (len(pattern), i)
        assert sum(t.draw_retries for t in o_traces) == sum(1 for r in pattern if r is rec)
