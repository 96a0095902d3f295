"""The reference's 12 dataset pairs as test fixtures (BASELINE configs[0]).

tests/golden/dataset_pack.npz holds, per pair, the JPEG file's bytes and the annotation read as gray
(what main.cpp reads: src/main.cpp:93 imread colour, :160-162 imread gray); it is written by
tests/golden/make_golden_inputs.py in the container that has /root/reference, so nothing here touches that tree.
"""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PACK = os.path.join(GOLD, "dataset_pack.npz")
NAMES = ["arara", "archespark", "dog", "flower", "heidelberg", "hills", "pigs", "rock", "straw", "streetart", "vintagegirl", "womanparasol"]
_pack = None


def have_pack():
    return os.path.exists(PACK)


def annotation_to_planes(bgr, ann):
    """ref: src/main.cpp:163-168 -- every annotation pixel != 32: edited BGR := that value, scribble := 255."""
    scribble = np.where(ann != 32, 255, 0).astype(np.uint8)
    edited = bgr.copy()
    edited[ann != 32] = ann[ann != 32][:, None]
    return scribble, edited


def load_pair(name):
    """-> (bgr, scribble, edited, annotation) of one dataset pair, decoded with the same cv2 the rest of the tests use."""
    global _pack
    import cv2
    if _pack is None:
        _pack = np.load(PACK)
    bgr = cv2.imdecode(_pack[name + "_jpg"], cv2.IMREAD_COLOR)
    ann = _pack[name + "_ann"]
    assert bgr is not None and bgr.shape[:2] == ann.shape
    scribble, edited = annotation_to_planes(bgr, ann)
    return bgr, scribble, edited, ann
