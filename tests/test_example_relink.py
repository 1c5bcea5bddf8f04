"""The drop-in claim, end to end: the reference's own example program (examples/04-disk-image-eqplane/disk-image.c, compiled
UNMODIFIED by oracle/Makefile `example` in the build container) linked against include/sim5lib.h + libsim5b200.so must print the
text dump of the build that links the reference's own library; and the batched entry + api.write_text_dump must
write that same dump.  /root/reference is not read here: both binaries travel in oracle/_ref/."""
import os
import subprocess

import numpy as np
import pytest

import harness as H
from sim5_b200 import abi

REF_BIN = os.path.join(H.ROOT, "oracle", "_ref", "disk-image-ref")
B200_BIN = os.path.join(H.ROOT, "oracle", "_ref", "disk-image-b200")
ARGS = ["0.9", "70"]          # the example's own usage line: <spin> <inclination>; 1280 x 720 pixels are compiled in


def _run(binary, timeout):
    r = subprocess.run([binary] + ARGS, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=timeout, text=True)
    assert r.returncode == 0, "%s exited with %d" % (binary, r.returncode)
    return r.stdout


def _parse(text):
    v = np.array(text.split(), dtype=np.float64).reshape(-1, 4)
    return v[:, 0].astype(np.int64), v[:, 1].astype(np.int64), v[:, 2], v[:, 3]


def _same_dump(a_txt, b_txt, label):
    """`%e` of a float keeps 7 digits: a 1e-15 difference between the doubles can move the last printed digit (1e-6 relative)."""
    la, lb = a_txt.split("\n"), b_txt.split("\n")
    assert len(la) == len(lb), "%s: %d lines vs %d" % (label, len(la), len(lb))
    same = sum(1 for x, y in zip(la, lb) if x == y)
    ya, xa, fa, ga = _parse(a_txt)
    yb, xb, fb, gb = _parse(b_txt)
    assert np.array_equal(ya, yb) and np.array_equal(xa, xb), label
    assert np.array_equal(fa == 0.0, fb == 0.0), "%s: different sets of lit pixels" % label
    for u, v, name in ((fa, fb, "flux"), (ga, gb, "g")):
        e = H.rel_err(u, v)
        assert e.max() <= 1.01e-6, "%s: %s differs by %.3e" % (label, name, e.max())
    frac = same / float(len(la))
    assert frac >= 0.9999, "%s: only %.6f of the lines are identical" % (label, frac)
    return frac


@pytest.mark.gpu
@pytest.mark.skipif(not (os.path.exists(REF_BIN) and os.path.exists(B200_BIN)), reason="oracle/_ref example binaries did not travel")
def test_example04_relinked_unmodified_and_text_dump(gpu_api, tmp_path):
    ref_txt = _run(REF_BIN, 120)
    # (1) the unmodified example on the scalar sim5lib.h API of libsim5b200.so: ~5 one-thread launches per pixel (~10 us each)
    got_txt = _run(B200_BIN, 900)
    f1 = _same_dump(got_txt, ref_txt, "relinked example")
    # (2) the same image through the batched entry and the text dump of the Python binding
    p = abi.default_params(1, 1280, 720)
    planes, st = gpu_api.trace_image(p, gpu_api.HostPlanes(p, pinned=False))
    path = os.path.join(str(tmp_path), "dump.txt")
    gpu_api.write_text_dump(path, planes, p)
    with open(path) as fh:
        f2 = _same_dump(fh.read(), ref_txt, "write_text_dump")
    print("example 04 (1280x720, a=0.9, i=70): relinked binary %.6f, batched dump %.6f of %d lines identical to the reference build's"
          % (f1, f2, len(ref_txt.split("\n"))))
