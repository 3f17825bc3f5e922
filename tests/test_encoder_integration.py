"""Drop-in proof (-m gpu): the reference encoder with xPatternSearch / xPatternSearchGT running on the
GPU (integration/_build/TAppEncoderHop) must write the SAME bitstream, reconstruction and syntax trace
as the unmodified CPU reference (oracle/_ref/TAppEncoderRef) on the same synthetic lenslet input."""
import hashlib
import os

import pytest

import _oracle
from hevc_hop_b200 import encoder

pytestmark = pytest.mark.gpu
have = os.path.exists(encoder.HOP_ENCODER) and os.path.exists(_oracle.REF_ENCODER)


@pytest.mark.skipif(not have, reason="encoder binaries not built (need /root/reference at build time)")
@pytest.mark.parametrize("size,bit_depth,qp", [((128, 128), 8, 32), ((128, 64), 8, 22), ((64, 64), 10, 32),
                                               ((136, 104), 8, 37)])      # partial CTUs at the right / bottom edge
def test_bitstream_identical_to_reference(size, bit_depth, qp):
    w, h = size
    ref = _oracle.encode_reference(w, h, seed=1, qp=qp, bit_depth=bit_depth)
    hop = encoder.encode(encoder.HOP_ENCODER, w, h, seed=1, qp=qp, bit_depth=bit_depth)
    assert hashlib.md5(hop["bitstream"]).hexdigest() == hashlib.md5(ref["bitstream"]).hexdigest()
    assert hop["rec"] == ref["rec"]
    assert hop["trace"] == ref["trace"]          # per-PU mvL0 / GT_FLAG / GT0L0..GT3L0 and bit counts
    assert len(hop["bitstream"]) > 100


@pytest.mark.skipif(not have, reason="encoder binaries not built (need /root/reference at build time)")
def test_unfused_call_path_is_identical_too(monkeypatch):
    """HOP_FUSED=0: xPatternSearch and xPatternSearchGT as separate GPU calls, fractional refinement on the host."""
    ref = _oracle.encode_reference(128, 64, seed=3)
    monkeypatch.setenv("HOP_FUSED", "0")
    hop = encoder.encode(encoder.HOP_ENCODER, 128, 64, seed=3)
    assert hop["bitstream"] == ref["bitstream"] and hop["rec"] == ref["rec"] and hop["trace"] == ref["trace"]


# ---- golden bitstreams of the CPU reference at the sizes BASELINE.json quotes --------------------------------
# tests/golden/encode_golden.json is written by tests/golden/make_encode_golden.py in the build container (the
# unmodified reference with the reference's OWN cfg files; 13 CPU-minutes for the 1024x1024 frame); here the GPU-backed
# encoder re-encodes the same frames with the repo's cfg files.  GPU runs get no retries.
import json  # noqa: E402

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encode_golden.json")
_golden = json.load(open(_GOLDEN)) if os.path.exists(_GOLDEN) else {}


def _encode_case(rec, **kw):
    args = dict(seed=rec["seed"], qp=rec["qp"], bit_depth=rec.get("bit_depth", 8), frames=rec.get("frames", 1))
    if rec.get("crop_of"):
        args["crop_of"] = tuple(rec["crop_of"])
    if rec.get("lowdelay_p"):
        args["cfg"] = encoder.CFG_LOWDELAY_P
    args.update(kw)
    return encoder.encode(encoder.HOP_ENCODER, rec["width"], rec["height"], **args)


@pytest.mark.skipif(not os.path.exists(encoder.HOP_ENCODER), reason="TAppEncoderHop not built")
@pytest.mark.parametrize("name", sorted(_golden))
def test_golden_bitstreams(name):
    """configs[0] (1024x1024 QP 32), a region of the configs[1] image at QP 22/27/32/37, Main10, and ISS + PSS frames:
    str.bin and rec.yuv byte-identical to the CPU reference's."""
    rec = _golden[name]
    hop = _encode_case(rec)
    assert len(hop["bitstream"]) == rec["bitstream_bytes"]
    assert hashlib.md5(hop["bitstream"]).hexdigest() == rec["bitstream_md5"]
    assert hashlib.md5(hop["rec"]).hexdigest() == rec["rec_md5"]


@pytest.mark.skipif(not have, reason="encoder binaries not built (need /root/reference at build time)")
def test_pss_slices_search_the_reset_mirror():
    """Low-delay P: the SS reference is reset to NOT_VALID for EVERY ISS / PSS slice (TComSlice.cpp:241-254); the
    mirror has to follow, or PSS searches would run on the previous picture's reconstruction."""
    ref = _oracle.encode_reference(128, 64, seed=4, frames=2, cfg=encoder.CFG_LOWDELAY_P)
    hop = encoder.encode(encoder.HOP_ENCODER, 128, 64, seed=4, frames=2, cfg=encoder.CFG_LOWDELAY_P, env_extra={"HOP_STATS": "1"})
    assert hop["bitstream"] == ref["bitstream"] and hop["rec"] == ref["rec"] and hop["trace"] == ref["trace"]
    assert "refReset+create" in hop["log"]


def _stat(log, key):
    import re
    m = re.search(key + r"\s+(\d+)", log)
    return int(m.group(1)) if m else None


@pytest.mark.skipif(not os.path.exists(encoder.HOP_ENCODER), reason="TAppEncoderHop not built")
def test_speculative_first_pu_searches_are_used_and_change_nothing():
    """the first PU of every partition mode of a CU is searched ahead of time.  With and without the
    speculation window the encoder writes the same bytes; with it, most first-PU requests are answered from the cache."""
    on = encoder.encode(encoder.HOP_ENCODER, 192, 128, seed=9, env_extra={"HOP_STATS": "1"})
    off = encoder.encode(encoder.HOP_ENCODER, 192, 128, seed=9, env_extra={"HOP_STATS": "1", "HOP_PREFETCH": "0"})
    noamp = encoder.encode(encoder.HOP_ENCODER, 192, 128, seed=9, env_extra={"HOP_STATS": "1", "HOP_PREFETCH_AMP": "0"})
    for other in (off, noamp):
        assert on["bitstream"] == other["bitstream"] and on["rec"] == other["rec"] and on["trace"] == other["trace"]
    calls, hits, misses = _stat(on["log"], "single-PU calls"), _stat(on["log"], "cache hits"), _stat(on["log"], "misses")
    assert calls == hits + misses and _stat(off["log"], "cache hits") == 0 and _stat(off["log"], "single-PU calls") == calls
    assert hits > 0.4 * calls, (calls, hits, misses)          # 18 k of 30.6 k calls are first PUs on a 512x512 image
    assert _stat(noamp["log"], "cache hits") <= hits
    cand = _stat(on["log"], "HOP candidates scored for the encoder")
    assert cand == _stat(off["log"], "HOP candidates scored for the encoder") and cand > 0


@pytest.mark.skipif(not os.path.exists(encoder.HOP_ENCODER_BATCH), reason="TAppEncoderHopBatch not built")
def test_long_lived_worker_writes_the_same_bytes():
    """integration/hop_batch_main.cpp: the reference's TAppEncTop run once per job inside one process that keeps the CUDA
    context (BASELINE configs[3]: a queue of images per GPU).  Every job must write exactly what a process of its own
    writes -- also the second time the same image comes by, and across bit depths / cfgs / sizes."""
    cases = [dict(width=128, height=128, seed=1, qp=32), dict(width=128, height=64, seed=3, qp=22),
             dict(width=64, height=64, seed=1, qp=32, bit_depth=10), dict(width=128, height=128, seed=1, qp=32),
             dict(width=128, height=64, seed=4, frames=2, cfg=encoder.CFG_LOWDELAY_P), dict(width=136, height=104, seed=1, qp=37)]
    w = encoder.EncoderWorker(env_extra={"HOP_STATS": "1"})
    try:
        outs = [w.encode(**c) for c in cases]
    finally:
        w.close()
    assert outs[0]["bitstream"] == outs[3]["bitstream"] and outs[0]["rec"] == outs[3]["rec"]
    for c, o in zip(cases, outs):
        one = encoder.encode(encoder.HOP_ENCODER, c["width"], c["height"], **{k: v for k, v in c.items() if k not in ("width", "height")})
        assert o["bitstream"] == one["bitstream"] and o["rec"] == one["rec"] and o["trace"] == one["trace"], c
        assert o["stats"] is not None and o["stats"]["gpu_search_calls"] > 0


@pytest.mark.skipif(not os.path.exists(encoder.HOP_ENCODER_BATCH), reason="TAppEncoderHopBatch not built")
def test_batch_driver_workers_and_processes_agree():
    from hevc_hop_b200 import batch
    tasks = [dict(width=128, height=64, seed=20 + i) for i in range(5)]
    a, _ = batch.encode_batch(tasks, procs=2, persistent=True)
    b, _ = batch.encode_batch(tasks, procs=2, persistent=False)
    assert all("error" not in r for r in a + b), (a, b)
    assert [r["md5"] for r in a] == [r["md5"] for r in b] and [r["rec_md5"] for r in a] == [r["rec_md5"] for r in b]
    assert sum(1 for r in a if r["worker_startup_s"] > 0) == 2
