"""Regenerates the golden fixtures under tests/golden/.

* hummingbird_gray_u8.npy -- the reference's test image resources/Hummingbird!.jpeg converted to 8-bit gray
  (PIL 'L'), the input of the reference's only exact end-to-end test (test/runtests.jl:167-175).  Needs
  /root/reference (present in the build container only).
* colony_gray_u8.npy, aged_whisky_gray_u8.npy -- the reference's SMALL and MED test images (resources/colony.png,
  resources/Aged Whisky.jpg) as 8-bit gray, inputs of its loss-bound tests (test/runtests.jl:72-140).
* config1_oracle.npz -- BASELINE config 1 (rand(50,100), ALL_METRICS, scales (1,2,4)) run through the CPU
  oracle.  The reference itself (Julia) cannot be executed in this image, so these vectors are oracle output,
  not Julia output; they pin the oracle against regressions and give the GPU tests a committed target.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import seq_oracle as O  # noqa: E402


def main():
    ref_img = "/root/reference/resources/Hummingbird!.jpeg"
    if os.path.exists(ref_img):
        from PIL import Image
        a = np.asarray(Image.open(ref_img).convert("L"), dtype=np.uint8)
        np.save(os.path.join(HERE, "hummingbird_gray_u8.npy"), a)
        for src, out in (("colony.png", "colony_gray_u8.npy"), ("Aged Whisky.jpg", "aged_whisky_gray_u8.npy")):
            a = np.asarray(Image.open(os.path.join(os.path.dirname(ref_img), src)).convert("L"), dtype=np.uint8)
            np.save(os.path.join(HERE, out), a)
    A = np.random.default_rng(2732).random((50, 100))
    r = O.sequence_oracle(A, (1, 2, 4), keep_matrices=False)
    np.savez_compressed(
        os.path.join(HERE, "config1_oracle.npz"), A=A, order=r.order, eta=r.eta, root=r.root,
        eta_groups=np.array([g.eta for g in r.groups]),
        eta_chunks=np.concatenate([np.array(g.eta_chunk) for g in r.groups]),
        roots_groups=np.array([g.root for g in r.groups]),
        parents=r.parents, mst_i=r.mst[0], mst_j=r.mst[1], mst_w=r.mst[2])


if __name__ == "__main__":
    main()
