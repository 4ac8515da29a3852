"""Result logs in the reference's formats (DefTracking.cc:321-328,507; GroundTruthFrame.cc:259-264;
GroundTruthCalculator.cc:174-186)."""
import numpy as np

from defslam_b200 import logs


def test_formats(tmp_path):
    L = logs.ResultLogs(str(tmp_path))
    L.frame(7.0, 412, 23, 980)
    L.frame(12345.9, 5, 0, 17)
    L.scale(7.0, 1.0340000391006470)      # float scale printed with 6 significant digits
    L.scale(1500000000.25, 0.5)
    name = L.errors(7.0, [0.00123456789, 12.5, 3.0, 1e-7])
    L.close()
    assert (tmp_path / "Matches.txt").read_text() == "00007 412 23 980\n12345 5 0 17\n"
    assert (tmp_path / "ScaleVariation.txt").read_text() == "7 1.034\n1.5e+09 0.5\n"
    assert name.endswith("ErrorGTs00007.txt")
    assert open(name).read() == "0.00123457\n      12.5\n         3\n     1e-07"
    # plotting.ipynb reads the error files with numpy.loadtxt
    assert np.allclose(np.loadtxt(name), [0.00123457, 12.5, 3.0, 1e-7])


def test_stream_writes_logs(tmp_path, oracle):
    from defslam_b200 import stream
    be = stream.Backend(oracle.load(), "oracle_", oracle.sft_solve)
    cfg = stream.StreamConfig(G=9, n_points=200, n_frames=6, nptsu=13, nptsv=15)
    stream.run_stream(be, cfg, logs_dir=str(tmp_path))
    lines = (tmp_path / "Matches.txt").read_text().splitlines()
    assert len(lines) == 6 and lines[0].split()[0] == "00000" and all(len(l.split()) == 4 for l in lines)
    assert (tmp_path / "ErrorGTs00005.txt").exists()
