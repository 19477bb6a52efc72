"""MOT-format readers / writer against the reference's parsing rules (src/data/mot17_dataset.cpp:149-289,
include/motcpp/utils/mot_format.hpp:20-74), and the replay tool end to end on the GPU."""
import os
import sys

import numpy as np
import pytest

from motcpp_b200 import mot_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_load_detections_mot17_comma_format(tmp_path):
    p = tmp_path / "det.txt"
    p.write_text("1,-1,100.5,200.25,50,80.5,0.9\n# comment\n\n1,-1,10,20,30,40,0.5,2\n3,-1,1,2,3\n2,-1,5,6,7,8,0.25,1,9,9\n")
    d = mot_io.load_detections(str(p))
    assert sorted(d) == [1, 2]                                                    # the 5-value row is skipped (:187)
    assert np.array_equal(d[1], np.array([[100.5, 200.25, 150.5, 280.75, 0.9, 0], [10, 20, 40, 60, 0.5, 2]], np.float32))
    assert np.array_equal(d[2], np.array([[5, 6, 12, 14, 0.25, 1]], np.float32))
    assert mot_io.load_detections(str(tmp_path / "missing.txt")) == {}


def test_load_detections_space_format_and_embeddings(tmp_path):
    p = tmp_path / "seq.txt"
    p.write_text("2 10 20 30 40 0.8 0\n1 1 2 3 4 0.7 1\n2 11 21 31 41 0.6 0\n")
    d = mot_io.load_detections(str(p))
    assert np.array_equal(d[2], np.array([[10, 20, 30, 40, 0.8, 0], [11, 21, 31, 41, 0.6, 0]], np.float32))
    assert np.array_equal(d[1], np.array([[1, 2, 3, 4, 0.7, 1]], np.float32))
    e = tmp_path / "emb.txt"
    e.write_text("0.1 0.2 0.3\n0.4 0.5 0.6\n0.7 0.8 0.9\n1 1 1\n")           # a 4th line beyond the detections is ignored
    m = mot_io.load_embeddings(str(e), d)
    assert np.array_equal(m[1], np.array([[0.1, 0.2, 0.3]], np.float32))          # ascending frame order
    assert np.array_equal(m[2], np.array([[0.4, 0.5, 0.6], [0.7, 0.8, 0.9]], np.float32))


def test_mot_format_writer(tmp_path):
    tracks = np.array([[100.7, 50.2, 180.9, 250.1, 3, 0.87654321, 0, 5], [-3.5, 10, 20.5, 40, 12, 1.0, 1, 0]], np.float32)
    m = mot_io.convert_to_mot_format(tracks, 7)
    assert m.shape == (2, 10) and np.all(m[:, 7:] == -1) and m[0, 1] == 3
    assert m[0, 4] == np.float32(180.9) - np.float32(100.7)
    txt = mot_io.format_mot_rows(m)
    # static_cast<int> truncates towards zero (x1 = -3.5 -> -3); conf fixed, 6 decimals (mot_format.hpp:62-73)
    assert txt == "7,3,100,50,80,199,0.876543,-1,-1,-1\n7,12,-3,10,24,30,1.000000,-1,-1,-1\n"
    out = tmp_path / "res" / "seq.txt"
    mot_io.write_mot_results(str(out), m)
    mot_io.write_mot_results(str(out), m[:1])                                     # appends (std::ios::app)
    assert out.read_text() == txt + txt.split("\n")[0] + "\n"
    assert mot_io.convert_to_mot_format(np.zeros((0, 8), np.float32), 1).shape == (0, 10)


def test_golden_fixture_equals_reader_when_reference_present():
    ref = "/root/reference/assets/MOT17-mini/train"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present (GPU box)")
    g = np.load(os.path.join(ROOT, "tests", "golden", "mot17_mini_dets.npz"))
    for seq in sorted(os.listdir(ref)):
        d = mot_io.load_detections(os.path.join(ref, seq, "det", "det.txt"))
        key = seq.replace("-", "_")
        frames, dets = g[key + "_frames"], g[key + "_dets"]
        for f in np.unique(frames):
            assert np.array_equal(d[int(f)], dets[frames == f]), (seq, f)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["sort", "bytetrack"])
def test_replay_tool_writes_what_the_oracle_tracks(oracle, tmp_path, method):
    from motcpp_b200 import _lib, build
    build.build()
    _lib.require_gpu()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import motb200_eval
    g = np.load(os.path.join(ROOT, "tests", "golden", "mot17_mini_dets.npz"))
    root = tmp_path / "train"
    seqs = sorted({k[:-len("_frames")] for k in g.files if k.endswith("_frames")})
    for key in seqs:                                         # write the fixture back out as MOT17 det.txt files
        os.makedirs(root / key / "det")
        with open(root / key / "det" / "det.txt", "w") as f:
            for fr, r in zip(g[key + "_frames"], g[key + "_dets"]):
                f.write(f"{fr},-1,{r[0]!r},{r[1]!r},{np.float32(r[2] - r[0])!r},{np.float32(r[3] - r[1])!r},{r[4]!r}\n".replace("np.float32(", "").replace(")", ""))
    out_dir = tmp_path / "results"
    assert motb200_eval.main(["motb200_eval", str(root), str(out_dir), method]) == 0
    for key in seqs:
        d = mot_io.load_detections(str(root / key / "det" / "det.txt"))
        ref = oracle.Sort(0.3, 1, 50, 3, 0.3) if method == "sort" else oracle.ByteTrack(0.3, 30, 50, 3, 0.3, 0.1, 0.45, 0.8, 30, 30)
        want = ""
        for fr in range(1, max(d) + 1):
            tr = ref.update(d.get(fr, np.zeros((0, 6), np.float32)))
            if fr in d:
                want += mot_io.format_mot_rows(mot_io.convert_to_mot_format(tr, fr))
        assert (out_dir / (key + ".txt")).read_text() == want, key
        assert len(want) > 1000
