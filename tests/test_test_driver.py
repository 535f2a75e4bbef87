"""The host-side mirror of the reference's test programs (eigensolver_gpu_b200/test_driver.py):
unformatted matrix files (test_driver/test_dsygvdx.F90:120-145) and the CPU-vs-custom comparison report."""
import numpy as np
import pytest

from eigensolver_gpu_b200 import test_driver as TD


@pytest.mark.parametrize("cplx", [False, True])
def test_unformatted_round_trip(tmp_path, cplx):
    a = TD.create_random_pd(37, cplx, seed=3)
    p = tmp_path / "a.bin"
    TD.write_unformatted(p, a, m=11, lda=37)
    raw = p.read_bytes()
    assert raw[:4] == (12).to_bytes(4, "little") and raw[16:20] == (12).to_bytes(4, "little")   # record markers
    b, n, m, lda = TD.read_unformatted(p, cplx)
    assert (n, m, lda) == (37, 11, 37)
    assert np.array_equal(a, b)
    with pytest.raises(ValueError):
        TD.read_unformatted(p, not cplx)          # wrong element type: record length does not match


def test_compare_metrics_match_the_reference_formulas():
    ref = np.array([1.0, 2.0, 0.0, -4.0])
    x = np.array([1.0, 2.2, 5.0, -4.0])
    l2, mx, i = TD.compare_1d(ref, x)             # entries with |ref| < 1e-10 are skipped (toolbox.F90)
    assert i == 1 and abs(mx - 10.0) < 1e-9
    assert abs(l2 - 0.2 / np.sqrt(21.0)) < 1e-12
    l2, mx, i, j = TD.compare_2d_abs(np.array([[1j, 2.0]]), np.array([[-1.0, 2.0]]))   # moduli: phases do not matter
    assert l2 == 0.0 and mx == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("cplx", [False, True])
def test_driver_program_from_files(tmp_path, cplx, capsys):
    from oracle import matgen
    n, m = 300, 120
    a, b = matgen.family_c(n, cplx, seed=21)
    TD.write_unformatted(tmp_path / "A.bin", a, m=m)
    TD.write_unformatted(tmp_path / "B.bin", b, m=m)
    rc = TD.main(["z" if cplx else "d", str(tmp_path / "A.bin"), str(tmp_path / "B.bin")])
    out = capsys.readouterr().out
    assert rc == 0
    assert "CPU_____________________" in out and "CUSTOM_____________________" in out
    assert "n,m,lda from files: 300 120 300" in out
    # the report lines carry the reference's metrics; parse the two l2norm errors (eigenvalues, |Z|)
    errs = [float(l.split()[2]) for l in out.splitlines() if l.strip().startswith("l2norm error")]
    assert len(errs) == 2 and errs[0] < 1e-13 and errs[1] < 1e-8
