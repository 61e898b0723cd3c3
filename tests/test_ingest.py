"""CPU tests of the native line-file readers (no GPU needed: host code of the C-ABI library)."""
import os

import numpy as np
import pytest

from radtxfr_b200 import ingest, synthetic as syn, _lib, build


@pytest.fixture(scope="module", autouse=True)
def _built():
    build.build()


def _par_line(M, I, nu, sw, a, ga, gs, el, n, d):
    # formats of HITRAN_DEFAULT_HEADER, misc/hapi.py:513-533
    def fx(fmt, v, w):                      # HITRAN drops the leading zero when the field would overflow
        t = fmt % v
        if len(t) > w:
            t = t.replace("0.", ".", 1)
        assert len(t) <= w, (fmt, v, t)
        return t.rjust(w)
    s = ("%2d%1d" % (M, I)) + fx("%12.6f", nu, 12) + fx("%10.3E", sw, 10) + fx("%10.3E", a, 10) + fx("%5.4f", ga, 5) \
        + fx("%5.3f", gs, 5) + fx("%10.4f", el, 10) + fx("%4.2f", n, 4) + fx("%8.6f", d, 8)
    return s + " " * (160 - len(s))


def test_par_reader_matches_python_fixed_width_parse(tmp_path):
    cols = syn.synthetic_lines(5, 300, 600.0, 2500.0, 1)
    rng = np.random.default_rng(0)
    rows = []
    for i in range(300):
        rows.append(_par_line(int(rng.integers(1, 40)), int(cols["local_iso_id"][i]), cols["nu"][i], cols["sw"][i],
                              rng.uniform(1e-3, 10), cols["gamma_air"][i], cols["gamma_self"][i], cols["elower"][i],
                              cols["n_air"][i], cols["delta_air"][i]))
    rows.insert(10, "garbage line that hapi would skip")
    rows.insert(20, " 1A" + rows[0][3:])                      # iso 'A': int() fails in hapi -> skipped
    rows.insert(30, "")                                       # blank
    p = tmp_path / "t.data"
    p.write_text("\n".join(rows) + "\n")
    got = ingest.read_par(p)
    exp = {"molec_id": [], "local_iso_id": [], "nu": [], "sw": [], "a": [], "gamma_air": [], "gamma_self": [],
           "elower": [], "n_air": [], "delta_air": []}
    sl = {"nu": (3, 15), "sw": (15, 25), "a": (25, 35), "gamma_air": (35, 40), "gamma_self": (40, 45),
          "elower": (45, 55), "n_air": (55, 59), "delta_air": (59, 67)}
    for r in rows:
        try:                                                    # hapi: int()/float() of the slices, skip on error
            rec = {"molec_id": int(r[0:2]), "local_iso_id": int(r[2:3])}
            for k, (a, b) in sl.items():
                rec[k] = float(r[a:b])
        except Exception:
            continue
        for k, v in rec.items():
            exp[k].append(v)
    assert len(got["nu"]) == len(exp["nu"]) == 300
    for k in exp:
        assert np.array_equal(np.asarray(got[k], dtype=float), np.asarray(exp[k], dtype=float)), k


def test_tape3_round_trip(tmp_path):
    t = syn.config2_tables(scale=0.005)
    cols = syn.merge_tables(t)
    p = tmp_path / "lines.tp3"
    ingest.write_tape3(p, cols)
    n = len(cols["nu"])
    assert os.path.getsize(p) == (4 + 1672 + 4) + ((n + 249) // 250) * (8 + 24 + 8 + 10000)
    back = ingest.read_tape3(p)
    assert len(back["nu"]) == n
    assert np.array_equal(back["nu"], cols["nu"]) and np.array_equal(back["molec_id"], cols["molec_id"])
    assert np.array_equal(back["local_iso_id"], cols["local_iso_id"])
    for k in ("gamma_air", "gamma_self", "elower", "n_air", "delta_air"):        # stored as float32
        assert np.allclose(back[k], cols[k], rtol=1e-6, atol=1e-7), k
    assert np.allclose(back["sw"], cols["sw"], rtol=2e-7)
    sub = ingest.read_tape3(p, 900.0, 1000.0)
    sel = (cols["nu"] >= 900.0) & (cols["nu"] <= 1000.0)
    assert np.array_equal(sub["nu"], cols["nu"][sel])
    by = ingest.split_by_molecule(back)
    assert sorted(by) == [1, 2, 3] and sum(len(v["nu"]) for v in by.values()) == n


def test_tape3_rejects_lfs_pointer(tmp_path):
    p = tmp_path / "AER-v3.6-0500-6000.tp3"
    p.write_text("version https://git-lfs.github.com/spec/v1\noid sha256:0000\nsize 970045456\n")
    with pytest.raises(_lib.RtxError, match="TAPE3"):
        ingest.read_tape3(p)
    with pytest.raises(_lib.RtxError, match="cannot open"):
        ingest.read_par(tmp_path / "missing.par")
