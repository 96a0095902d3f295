"""Run HERE (the container that has /root/reference): packs one dataset pair into
tests/golden/inputs_<name>.npz so that GPU-box runs never need /root/reference.

ref: src/main.cpp:93 (imread colour), :160-170 (annotation: gray read, 32 = unannotated).
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/dataset"


def main(names):
    for name in names:
        bgr = cv2.imread(os.path.join(REF, "images", name + ".jpg"))
        ann = cv2.imread(os.path.join(REF, "annotations", name + ".png"), 0)
        assert bgr is not None and ann is not None and bgr.shape[:2] == ann.shape
        out = os.path.join(ROOT, "tests", "golden", "inputs_%s.npz" % name.lower())
        np.savez_compressed(out, bgr=bgr, annotation=ann)
        print(out, bgr.shape, os.path.getsize(out))


if __name__ == "__main__":
    main(sys.argv[1:] or ["Dog"])
