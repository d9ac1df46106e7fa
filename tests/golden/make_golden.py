"""Generates the committed golden fixtures (run in the build container, where
/root/reference is mounted; the GPU box only reads the resulting files).

  testpeaq_vectors.npz  the 12 golden arrays of the reference's own unit test
                        (/root/reference/src/testpeaq.c:37-599), parsed from
                        the C source as plain numbers.
  ref_outputs.npz       outputs of the REFERENCE ITSELF (oracle/_ref/libpeaq_ref.so,
                        i.e. /root/reference/src/*.c compiled by oracle/Makefile)
                        on seeded inputs: known-answer signals of
                        runtest-1.0.sh and synthetic / noise pairs, basic and
                        advanced mode: MOVs, DI, ODG, frame counts.

usage: make -C oracle ref && python tests/golden/make_golden.py
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

from refharness import RefPeaq, audiotestsrc, as_interleaved  # noqa: E402
from signals import golden_cases  # noqa: E402


def parse_testpeaq(path):
    src = open(path).read()
    out = {}
    for m in re.finditer(r"static\s+(?:const\s+)?g?double\s+(\w+)\s*\[\s*\]\s*=\s*\{(.*?)\};", src, re.S):
        name, body = m.group(1), m.group(2)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        vals = [float(x) for x in re.findall(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?", body)]
        out[name] = np.array(vals, dtype=np.float64)
    return out


def main():
    vec = parse_testpeaq("/root/reference/src/testpeaq.c")
    for k, v in vec.items():
        print(k, v.shape)
    np.savez_compressed(os.path.join(HERE, "testpeaq_vectors.npz"), **vec)

    out = {}
    for name, (ref, test, channels) in golden_cases().items():
        for adv in (0, 1):
            r = RefPeaq(bool(adv), 92.0, channels).run(ref, test)
            key = "%s|%s" % (name, "advanced" if adv else "basic")
            movs = np.full(11, np.nan)
            movs[:len(r["movs"])] = r["movs"]
            out[key] = np.concatenate([[r["odg"], r["di"], r["totalsnr"], r["frames_fft"], r["frames_fb"],
                                        float(r["loudness_reached_frame"]), len(r["movs"])], movs])
            print(key, "odg %.6f di %.6f" % (r["odg"], r["di"]))
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)


if __name__ == "__main__":
    main()
