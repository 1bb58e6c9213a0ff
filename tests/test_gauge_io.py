"""Gauge configurations in the reference's file formats (csrc/gauge_io.cu).

Host side (no GPU): the C++ reader / writer against (i) an independent Python writer in this file, (ii) the oracle's Python
reader, (iii) the golden link arrays under tests/golden (made from the reference's fixtures by tests/golden/make_golden.py) and,
when /root/reference is present (build container only), (iv) the reference's OWN files test/confs_*/conf_00000100.ildg{,.txt}:
reading them gives the golden arrays, and writing the golden arrays reproduces them BYTE FOR BYTE -- LIME header, big-endian
payload, and Julia's shortest-round-trip number formatting in the text format.  This is one of the few places where the
reference's own data pins the new code exactly.
Device side (staged GPU test, pre-flighted under tests/emu): file -> device links -> file round trips and D built on loaded links."""
import struct
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "latticeqcd.jl_b200"))
import lqcd_b200 as q                       # noqa: E402
from oracle import oracle as orc            # noqa: E402

REF = Path("/root/reference/test")
FIXTURES = {          # golden array -> the reference directory it was made from (tests/golden/fixtures.json)
    "wilson_4444": "confs_HMC_L04040404_beta5.7_Wilson_kappa0.141139",
    "staggered_4444": "confs_HMC_L04040404_beta5.7_Staggered_mass0.5",
    "staggered_nf2_4444": "confs_HMC_L04040404_beta5.7_Staggered_mass0.5_Nf2",
    "quenched_su3_4444": "confs_HMC_L04040404_beta5.7_quenched_su3",
}
DIMS = (4, 4, 4, 4)


def py_write_ildg(path, U):
    """independent restatement of the format: U [mu,t,z,y,x,b,a] -> LIME record of big-endian doubles"""
    f = np.ascontiguousarray(U.transpose(1, 2, 3, 4, 0, 6, 5))                 # [t,z,y,x,mu,a,b]
    payload = np.stack([f.real, f.imag], axis=-1).astype(">f8").tobytes()
    hdr = struct.pack(">IHHQ", 0x456789AB, 1, 0xC000, len(payload)) + b"ildg-binary-data".ljust(128, b"\0")
    Path(path).write_bytes(hdr + payload)


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_reader_and_writer_against_python_restatements(tmp_path, golden_dir, name):
    U = np.load(golden_dir / f"{name}.npy")
    py_write_ildg(tmp_path / "a.ildg", U)
    got = q.load_gaugefield(tmp_path / "a.ildg", DIMS, "ILDG")
    assert np.array_equal(got.data, U)
    q.save_binarydata(got, tmp_path / "b.ildg")
    assert (tmp_path / "b.ildg").read_bytes() == (tmp_path / "a.ildg").read_bytes()
    assert np.array_equal(orc.load_ildg(tmp_path / "b.ildg", DIMS), U)
    q.save_textdata(got, tmp_path / "b.txt")
    assert np.array_equal(orc.load_bridgetext(tmp_path / "b.txt", DIMS), U)      # shortest digits round-trip exactly
    U2 = q.Initialize_Gaugefields(3, 0, *DIMS, condition="cold")
    q.load_BridgeText_(tmp_path / "b.txt", U2)
    assert np.array_equal(U2.data, U)


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_reference_files_byte_for_byte(tmp_path, golden_dir, name):
    U = np.load(golden_dir / f"{name}.npy")
    ref_bin, ref_txt = REF / FIXTURES[name] / "conf_00000100.ildg", REF / FIXTURES[name] / "conf_00000100.ildg.txt"
    assert np.array_equal(q.load_gaugefield(ref_bin, DIMS, "ILDG").data, U)
    assert np.array_equal(q.load_gaugefield(ref_txt, DIMS, "BridgeText").data, U)
    q.save_binarydata(U, tmp_path / "w.ildg")
    assert (tmp_path / "w.ildg").read_bytes() == ref_bin.read_bytes()
    q.save_textdata(U, tmp_path / "w.txt")
    assert (tmp_path / "w.txt").read_bytes() == ref_txt.read_bytes()


@pytest.mark.skipif(not REF.exists(), reason="the reference tree is only present in the build container")
@pytest.mark.parametrize("d,nc,dims", [("confs_HMC_L04040404_beta2.5_quenched_su2", 2, (4, 4, 4, 4)), ("confs_HMC_L04040404_beta5.7_quenched_su4", 4, (4, 4, 4, 4)),
                                       ("confs_HMC_L04040404_beta5.7_Domainwall", 3, (4, 4, 2, 2))])
def test_other_gauge_groups_and_shapes(tmp_path, d, nc, dims):
    """the host reader / writer is NC-generic (the reference ships SU(2), SU(4) and 4.4.2.2 fixtures): text == binary, unitary,
    and written files are byte-identical"""
    cands = [p for p in REF.glob(d.replace("beta2.5", "beta*").replace("beta5.7", "beta*")) if (p / "conf_00000100.ildg").exists()]
    if not cands:
        pytest.skip("fixture directory not found")
    p = cands[0]
    A = q.load_gaugefield(p / "conf_00000100.ildg", dims, "ILDG", NC=nc)
    B = q.load_gaugefield(p / "conf_00000100.ildg.txt", dims, "BridgeText", NC=nc)
    A, B = (A.data if nc == 3 else A), (B.data if nc == 3 else B)
    assert np.array_equal(A, B)
    M = np.swapaxes(A, -1, -2)
    assert np.abs(np.einsum("...ij,...kj->...ik", M, M.conj()) - np.eye(nc)).max() < 1e-9
    q.save_binarydata(A, tmp_path / "w.ildg")
    assert (tmp_path / "w.ildg").read_bytes() == (p / "conf_00000100.ildg").read_bytes()
    q.save_textdata(A, tmp_path / "w.txt")
    assert (tmp_path / "w.txt").read_bytes() == (p / "conf_00000100.ildg.txt").read_bytes()


def test_julia_number_formatting(tmp_path):
    """save_textdata prints like Julia's print(::Float64): positional while the decimal point sits in (-4, 6], else d.ddde<exp>"""
    vals = [1.0, -1.0, 0.0, 0.5, 0.1, 1e-4, 1.5e-4, 1e-5, 7.932689995028509e-5, -2.5e-7, 123456.0, 1234567.0, 1e6, 1e22, 0.30000000000000004,
            5e-324, 1.7976931348623157e308, 100.0, 123.456, -0.0]
    want = ["1.0", "-1.0", "0.0", "0.5", "0.1", "0.0001", "0.00015", "1.0e-5", "7.932689995028509e-5", "-2.5e-7", "123456.0", "1.234567e6", "1.0e6", "1.0e22",
            "0.30000000000000004", "5.0e-324", "1.7976931348623157e308", "100.0", "123.456", "-0.0"]
    U = np.zeros((4, 1, 1, 1, 1, 3, 3), dtype=complex)
    flat = np.zeros(72)
    flat[: len(vals)] = vals
    f = np.zeros(36, dtype=complex)
    f.real, f.imag = flat[0::2], flat[1::2]                       # (not re + 1j*im: that loses the sign of -0.0)
    f = f.reshape(4, 3, 3)                                        # file order [mu, a, b]
    U[:, 0, 0, 0, 0] = np.swapaxes(f, -1, -2)
    q.save_textdata(U, tmp_path / "n.txt")
    lines = (tmp_path / "n.txt").read_text().split("\n")
    assert lines[: len(vals)] == want


def test_errors_are_reported(tmp_path, golden_dir):
    U = np.load(golden_dir / "wilson_4444.npy")
    with pytest.raises(q.LqcdError, match="cannot open"):
        q.load_gaugefield(tmp_path / "missing.ildg", DIMS)
    py_write_ildg(tmp_path / "a.ildg", U)
    with pytest.raises(q.LqcdError, match="does not match"):
        q.load_gaugefield(tmp_path / "a.ildg", (4, 4, 4, 8))
    (tmp_path / "junk.ildg").write_bytes(b"\0" * 400)
    with pytest.raises(q.LqcdError, match="not a LIME file"):
        q.load_gaugefield(tmp_path / "junk.ildg", DIMS)
    q.save_textdata(U, tmp_path / "a.txt")
    txt = (tmp_path / "a.txt").read_text().split("\n")
    (tmp_path / "short.txt").write_text("\n".join(txt[:1000]) + "\n")
    with pytest.raises(q.LqcdError, match="file ends"):
        q.load_gaugefield(tmp_path / "short.txt", DIMS, "BridgeText")
    (tmp_path / "bad.txt").write_text("\n".join(txt[:10] + ["abc"] + txt[11:]))
    with pytest.raises(q.LqcdError, match="not a number"):
        q.load_gaugefield(tmp_path / "bad.txt", DIMS, "BridgeText")


# ---- device ----------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(4, 4, 4, 4), (8, 4, 6, 4), (32, 4, 2, 2)])
@pytest.mark.parametrize("fmt", ["ILDG", "BridgeText"])
def test_device_load_and_save(tmp_path, dims, fmt):
    U = orc.random_su3(dims, seed=5, eps=0.4)
    (q.save_binarydata if fmt == "ILDG" else q.save_textdata)(U, tmp_path / "in")
    ctx = q.get_context(dims)
    q.load_gaugefield_device_(ctx, tmp_path / "in", fmt)
    assert np.array_equal(q.get_links(ctx), U)
    assert abs(q.plaquette(ctx) - orc.plaquette(dims, U)) < 1e-12
    q.save_gaugefield_device(ctx, tmp_path / "out", fmt)
    assert (tmp_path / "out").read_bytes() == (tmp_path / "in").read_bytes()
    # the operator on loaded links == the oracle on the same links
    Ug = q.gaugefields_from_array(U)
    x = q.Initialize_pseudofermion_fields(Ug[0], "Wilson")
    D = q.Dirac_operator(Ug, x, {"Dirac_operator": "Wilson", "κ": 0.12, "boundarycondition": [1, 1, 1, -1]})
    q.load_gaugefield_device_(ctx, tmp_path / "in", fmt)             # D is bound to the context's device links
    src = orc.gaussian_field(dims, orc.WILSON, seed=8)
    y = q.similar(x)
    q.mul_(y, D, x.from_host(src))
    want = orc.apply(orc.make_op(dims, kappa=0.12), orc.WILSON, orc.D, U, src)
    assert np.abs(y.to_host() - want).max() / np.abs(want).max() < 1e-13
