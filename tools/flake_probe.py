"""Which side varies?  Encode the partial-CTU integration case repeatedly with the unmodified CPU reference and the
GPU-backed encoder and count the distinct bitstreams each one produces:  python tools/flake_probe.py [runs]"""
import collections, hashlib, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import conftest  # noqa: F401
import _oracle
from hevc_hop_b200 import encoder

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 30
refs, hops = collections.Counter(), collections.Counter()
for i in range(runs):
    try:
        r = _oracle.encode_reference(136, 104, seed=1, qp=37, bit_depth=8)
        refs[hashlib.md5(r["bitstream"] + r["rec"] + r["trace"]).hexdigest()[:10]] += 1
    except Exception as e:
        refs["ERR " + str(e)[:60]] += 1
    try:
        h = encoder.encode(encoder.HOP_ENCODER, 136, 104, seed=1, qp=37, bit_depth=8)
        hops[hashlib.md5(h["bitstream"] + h["rec"] + h["trace"]).hexdigest()[:10]] += 1
    except Exception as e:
        hops["ERR " + str(e)[:60]] += 1
print("reference:", dict(refs))
print("gpu      :", dict(hops))
