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
