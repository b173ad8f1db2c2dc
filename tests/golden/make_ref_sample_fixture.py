#!/usr/bin/env python3
"""Derives a small statistical fixture from the ONLY outputs the reference publishes: its two
sample renders of Asset/cbox.json (1024x768, ACES, RGBA8),

    /root/reference/Sample - Path Tracing.png
    /root/reference/Sample - Primary Sample Space Metropolis Light Transport.png

    python tests/golden/make_ref_sample_fixture.py        # build container only (needs /root/reference)

writes tests/golden/ref_sample_blocks16.npz:
    pt, mlt        float32 [48, 64, 3]   mean 8-bit RGB of every 16x16-pixel block
    pt_black, mlt_black   number of exactly-black pixels (the reference's NaN pixels)

The images are tone-mapped 8-bit renders at an unknown sample count, so this cannot pin bits; it pins
the *converged picture* (camera, geometry, light power, BSDFs, MIS weights, ACES + gamma) that the F#
program produces — see tests/test_ref_sample_image.py for what is compared and what is masked.
"""
import os

import numpy as np
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
B = 16


def blocks(path):
    a = np.asarray(Image.open(path)).astype(np.float64)[..., :3]
    h, w, _ = a.shape
    return a.reshape(h // B, B, w // B, B, 3).mean(axis=(1, 3)).astype(np.float32), int((a.sum(axis=2) == 0).sum())


def main():
    pt, ptb = blocks(os.path.join(REF, "Sample - Path Tracing.png"))
    mlt, mltb = blocks(os.path.join(REF, "Sample - Primary Sample Space Metropolis Light Transport.png"))
    np.savez_compressed(os.path.join(HERE, "ref_sample_blocks16.npz"), pt=pt, mlt=mlt, pt_black=ptb, mlt_black=mltb)
    print(pt.shape, mlt.shape, ptb, mltb)


if __name__ == "__main__":
    main()
