"""CPU tests of the oracle itself: known answers from SURVEY.md Appendix A, the OpenCV
restatements against cv2, and the golden vectors recorded from the reference's own kernels
on a B200 (tests/golden/ref_*.npz, tests/golden/gen_golden_ref.py)."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import binding as ob
from realtimedepthdiffusion_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_lut_known_answers():
    lut = ob.load_weights(0.4)
    assert lut[0] == 1.0 and lut[256] == 0.0
    want = np.exp((np.float32(-0.4) * np.arange(256, dtype=np.float32)).astype(np.float64)).astype(np.float32)
    assert (np.abs(lut[:219] - want[:219]) <= np.spacing(want[:219])).all() and lut[1] == np.float32(0.67032003)
    tiny = np.finfo(np.float32).tiny
    assert lut[218] >= tiny and (lut[219:256] < tiny).all() and (lut[219:256] > 0).all()   # denormals kept
    assert lut[255] == np.float32(6e-45)


def test_omega_schedule_known_answers():
    om = ob.omega_schedule(100)
    assert (om[:10] == 1.0).all()
    assert om[10] == np.float32(1.9609766) and om[11] == np.float32(1.9248844)
    assert om[30] == np.float32(1.7538558) and (om[61:] == np.float32(1.7527453)).all()


def test_float_to_u8_gate_semantics():
    # trunc toward zero, negatives -> 0, >= 256 keeps the low byte (300.7 -> 44, 256 -> 0)
    gray = np.array([[10, 200]], np.uint8)
    for d0, d1, lvl, expect in ((3.9, 8.2, 1, 190), (3.9, 7.9, 1, 0), (0.2, 1.0, 0, 190), (0.2, 0.9, 0, 0),
                                (-3.0, 0.5, 0, 0), (300.7, 40.0, 1, 0), (256.0, 5.0, 1, 190)):
        idx = ob.index_to_weight(gray, np.array([[d0, d1]], np.float32), lvl, 5)
        assert idx[0, 0].tolist() == [256, expect, 256, 256], (d0, d1, lvl)
        assert idx[0, 1].tolist() == [expect, 256, 256, 256]
    idx = ob.index_to_weight(gray, np.zeros((1, 2), np.float32), 5, 5)     # coarsest: ungated
    assert idx[0, 0, 1] == 190


def test_single_pixel_and_all_scribble():
    lut = ob.load_weights()
    d = ob.solve_level(np.array([[77.0]], np.float32), np.zeros((1, 1), np.uint8), np.array([[5]], np.uint8), 3, 0, 0, lut)
    # no neighbours: count == 0 -> result 0; x1 = .99*(0-77)+77 = 0.77..., relaxes toward 0
    assert 0.0 <= d[0, 0] < 1.0
    x = np.random.default_rng(0).uniform(0, 255, (9, 13)).astype(np.float32)
    d = ob.solve_level(x, np.full((9, 13), 255, np.uint8), np.zeros((9, 13), np.uint8), 40, 0, 1, lut)
    assert (d == x).all()


def test_one_sweep_by_hand():
    lut = ob.load_weights()
    gray = np.array([[0, 1, 3]], np.uint8)
    x = np.array([[10.0, 20.0, 60.0]], np.float32)
    s = np.array([[255, 0, 255]], np.uint8)
    d = ob.solve_level(x, s, gray, 1, 0, 0, lut)   # level == maxLevel: ungated
    w1, w2 = lut[1], lut[2]
    sm = np.float32(np.float32(w1 * np.float32(10.0)) )       # fma(w1,10,0)
    import math
    sm = np.float32(math.fma(float(w2), 60.0, float(sm))) if hasattr(math, "fma") else None
    if sm is not None:
        r = np.float32(sm / np.float32(w1 + w2))
        t = np.float32(r - np.float32(20.0))
        u = np.float32(math.fma(float(np.float32(0.99)), float(t), 20.0))
        out = np.float32(math.fma(1.0, float(u), 0.0))
        assert d[0, 1] == out
    assert d[0, 0] == 10.0 and d[0, 2] == 60.0


def test_opencv_restatements_match_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for rows, cols in ((67, 120), (135, 241), (1, 9), (8, 1), (50, 51)):
        bgr = rng.integers(0, 256, (rows, cols, 3), dtype=np.uint8)
        assert (ob.bgr2gray(bgr) == cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)).all()
        g = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
        if rows > 2 and cols > 2:
            assert (ob.pyrdown_gray(g) == cv2.pyrDown(g)).all()
        f = rng.uniform(-5, 260, (rows, cols)).astype(np.float32)
        q = np.empty((rows, cols), np.uint8)
        cv2.convertScaleAbs  # noqa: B018  (presence check only)
        assert (ob.quantise_u8(f) == np.clip(np.rint(f), 0, 255).astype(np.uint8)).all()
        if rows > 1 and cols > 1:
            for dr, dc in ((2 * rows, 2 * cols), (2 * rows + 1, 2 * cols + 1), (2 * rows, 2 * cols + 1)):
                want = cv2.pyrUp(f, dstsize=(dc, dr))
                got = ob.pyrup_f32(f, dr, dc)
                # bit-equal to this image's cv2 build; allow 2 ulp for builds whose v_muladd fuses (see oracle header)
                assert np.allclose(got, want, rtol=3e-7, atol=1e-5), (rows, cols, dr, dc, np.abs(got - want).max())


def test_image_ops_semantics():
    rng = np.random.default_rng(5)
    ps = (rng.random((11, 14)) < 0.2).astype(np.uint8) * 255
    pe = rng.integers(0, 256, (11, 14, 3), dtype=np.uint8)
    cs0 = np.zeros((5, 7), np.uint8)
    ce0 = np.full((5, 7, 3), 9, np.uint8)
    cs, ce = ob.pyrdown_annotation(ps, pe, cs0, ce0)
    for y in range(5):
        for x in range(7):
            hit = None
            for py in (2 * y - 1, 2 * y):
                for px in (2 * x - 1, 2 * x):
                    if 0 <= py < 11 and 0 <= px < 14 and ps[py, px] == 255:
                        hit = (py, px)
            if hit is None:
                assert cs[y, x] == 0 and (ce[y, x] == 9).all()
            else:
                assert cs[y, x] == 255 and ce[y, x, 0] == pe[hit[0], hit[1], 0] and (ce[y, x, 1:] == 9).all()
    e, s = ob.paint(3, 2, 128, 5, np.zeros((8, 9, 3), np.uint8), np.zeros((8, 9), np.uint8))
    want = np.zeros((8, 9), np.uint8)
    want[0:5, 1:6] = 255                       # half side = 5 // 2 = 2
    assert (s == want).all() and (e[..., 1] == want // 255 * 128).all()
    bgr, scr, ed = synth.synth_case(40, 60, 1)
    s2, e2 = np.zeros_like(scr), bgr.copy()
    for ev in synth.brush_events(40, 60, 1, 2, 24):
        e2, s2 = ob.paint(*ev, e2, s2)
    # synth.paint_events is the vectorised twin of oracle_paint
    sc, edd = synth.paint_events(bgr, synth.brush_events(40, 60, 1, 2, 24))
    assert (sc == s2).all() and (edd == e2).all()


def test_defocus_kernel_size_table():
    assert ob.defocus_kernel_size(1080, 1920) == 55
    assert ob.defocus_kernel_size(2160, 3840) == 110
    assert ob.defocus_kernel_size(16384, 16384) == 579


def _golden(pattern):
    files = sorted(glob.glob(os.path.join(GOLD, pattern)))
    if not files:
        pytest.skip("no golden vectors %s (generate with tests/golden/gen_golden_ref.py on a GPU box)" % pattern)
    return files


def _load_case(name):
    if name in ("dog", "womanparasol"):
        z = np.load(os.path.join(GOLD, "inputs_%s.npz" % name))
        bgr, ann = z["bgr"], z["annotation"]
        scribble = np.where(ann != 32, 255, 0).astype(np.uint8)
        edited = bgr.copy()
        edited[ann != 32] = ann[ann != 32][:, None]
        return bgr, scribble, edited
    cases = {"synth_odd": (203, 317, 11), "synth_small": (96, 130, 12), "synth_tiny": (45, 47, 13)}
    return synth.synth_case(*cases[name])


@pytest.mark.parametrize("name", ["synth_tiny", "synth_small", "synth_odd", "dog", "womanparasol"])
def test_oracle_reproduces_reference_solver_golden(name):
    path = os.path.join(GOLD, "ref_solver_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden %s missing" % path)
    z = np.load(path)
    bgr, scribble, edited = _load_case(name)
    st = ob.FrameState(bgr)
    assert st.levels == int(z["levels"])
    u8 = st.solve(scribble, edited, int(z["max_iterations"]), keep_levels=True)
    for l in range(st.levels - 1, -1, -1):
        assert sha(st.per_level[l]["in"]) == str(z["in_sha_%d" % l]), "level %d input" % l
        if "out_%d" % l in z:
            assert (st.per_level[l]["out"] == z["out_%d" % l]).all(), "level %d" % l
        assert sha(st.per_level[l]["out"]) == str(z["out_sha_%d" % l]), "level %d output" % l
    assert (u8 == z["depth_u8"]).all()
    ev = synth.brush_events(st.rows, st.cols, 99, 1, 6)
    s2, e2 = synth.paint_events(bgr, ev, scribble.copy(), edited.copy())
    u8b = st.solve(s2, e2, int(z["max_iterations"]), keep_levels=True)
    assert sha(st.per_level[0]["out"]) == str(z["frame2_out_sha_0"])
    assert (u8b == z["frame2_depth_u8"]).all()


def _dataset_golden():
    import json
    path = os.path.join(GOLD, "ref_dataset.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/ref_dataset.json missing (tests/golden/gen_golden_ref.py --dataset on a GPU box)")
    return json.load(open(path))


@pytest.mark.parametrize("name", ["arara", "archespark", "dog", "flower", "heidelberg", "hills", "pigs", "rock", "straw", "streetart",
                                  "vintagegirl", "womanparasol"])
def test_oracle_reproduces_reference_on_every_dataset_pair(name):
    """BASELINE configs[0], all 12 pairs: 4- and 5-level pyramids, ten of them with floor/ceil size mismatches between levels
    (SURVEY.md Appendix B); the hashes were recorded from the reference's own kernels on a B200."""
    from tests import dataset
    gold = _dataset_golden()[name]
    bgr, scribble, edited, _ = dataset.load_pair(name)
    st = ob.FrameState(bgr)
    assert st.levels == gold["levels"] and (st.rows, st.cols) == (gold["rows"], gold["cols"])
    u8 = st.solve(scribble, edited, 1000, keep_levels=True)
    for l in range(st.levels - 1, -1, -1):
        assert sha(st.per_level[l]["in"]) == gold["in_sha"][str(l)], "level %d input" % l
        assert sha(st.per_level[l]["out"]) == gold["out_sha"][str(l)], "level %d output" % l
    assert sha(u8) == gold["depth_u8_sha"]
    ev = synth.brush_events(st.rows, st.cols, 99, 1, 6)
    s2, e2 = synth.paint_events(bgr, ev, scribble.copy(), edited.copy())
    u8b = st.solve(s2, e2, 1000, keep_levels=True)
    assert sha(st.per_level[0]["out"]) == gold["frame2_out_sha_0"]
    assert sha(u8b) == gold["frame2_depth_u8_sha"]


def test_oracle_reproduces_reference_weights_golden():
    for f in _golden("ref_weights_*.npz"):
        z = np.load(f)
        idx = ob.index_to_weight(z["gray"], z["depth"], int(z["level"]), int(z["levels"]) - 1)
        packed = np.stack([idx[..., 0] * 1000 + idx[..., 1], idx[..., 2] * 1000 + idx[..., 3]], -1)
        assert (packed == z["int2"]).all(), f


def test_oracle_reproduces_reference_effects_golden():
    for f in _golden("ref_effects_synth_*.npz"):
        z = np.load(f)
        name = os.path.basename(f)[len("ref_effects_"):-4]
        bgr, _, _ = _load_case(name)
        gray = ob.bgr2gray(bgr)
        depth = z["depth"]
        assert (ob.desaturate(bgr, gray, depth) == z["GPUSimulateDesaturation"]).all()
        assert (ob.defocus(bgr, depth) == z["GPUSimulateDefocus"]).all()
        # haze: libdevice expf vs glibc expf -- at most one grey level on a tiny fraction of values
        h = ob.haze(bgr, depth).astype(np.int16) - z["GPUSimulateHaze"].astype(np.int16)
        assert np.abs(h).max() <= 1 and (h != 0).mean() < 1e-3


def test_oracle_reproduces_reference_effects_on_dog():
    """Config 1 end to end on the CPU: oracle solve of Dog (bit-exact, sha-pinned) then the three effects."""
    path = os.path.join(GOLD, "ref_effects_dog.npz")
    if not os.path.exists(path):
        pytest.skip("golden missing")
    z = np.load(path)
    bgr, scribble, edited = _load_case("dog")
    st = ob.FrameState(bgr)
    st.solve(scribble, edited, 1000)
    st.solve(*synth.paint_events(bgr, synth.brush_events(st.rows, st.cols, 99, 1, 6), scribble.copy(), edited.copy()), 1000)
    depth = st.depth[0]
    assert sha(depth) == str(z["depth_sha"])
    gray = ob.bgr2gray(bgr)
    assert sha(ob.desaturate(bgr, gray, depth)) == str(z["desaturation_sha"])
    assert sha(ob.defocus(bgr, depth)) == str(z["defocus_sha"])
    h = ob.haze(bgr, depth)[::4, ::4].astype(np.int16) - z["haze_sample"].astype(np.int16)
    assert np.abs(h).max() <= 1 and (h != 0).mean() < 1e-3
