"""The C++ host side (adapter/DefOptimizerB200.h, same signature and side effects as the reference's
Optimizer::DefPoseOptimization) compiled against mock DefSLAM types and run end to end."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "_emu", "test_adapter")


EXE2 = os.path.join(ROOT, "tests", "_emu", "test_adapter_nrsfm")


def _build(oracle, src="test_adapter.cc", exe=EXE):
    oracle.load()
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    cmd = ["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", src),
           "-L" + os.path.join(ROOT, "defslam_b200"), "-ldefslam_b200",
           "-L" + os.path.join(ROOT, "oracle", "_build"), "-loracle",
           "-Wl,-rpath," + os.path.join(ROOT, "defslam_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle", "_build")]
    subprocess.run(cmd, check=True, capture_output=True)


def _run(exe=EXE):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    return r.returncode, r.stdout + r.stderr


def test_adapter_compiles_and_fails_loudly_without_a_device(oracle, cuda_lib):
    _build(oracle)
    if cuda_lib.defslam_device_count() > 0:
        pytest.skip("device present: covered by the gpu test")
    rc, out = _run()
    assert rc == 0, out
    assert "untouched" in out


@pytest.mark.gpu
def test_adapter_matches_oracle_on_gpu(oracle, cuda_lib):
    _build(oracle)
    rc, out = _run()
    assert rc == 0, out
    assert "max node err" in out


def test_nrsfm_adapter_compiles_and_fails_loudly_without_a_device(oracle, cuda_lib):
    """adapter/NrsfmB200.h: calculateSchwarps / ObtainK1K2 / estimateSurface against mock DefSLAM types."""
    _build(oracle, "test_adapter_nrsfm.cc", EXE2)
    if cuda_lib.defslam_device_count() > 0:
        pytest.skip("device present: covered by the gpu test")
    rc, out = _run(EXE2)
    assert rc == 0, out
    assert "untouched" in out


@pytest.mark.gpu
def test_nrsfm_adapter_matches_oracle_on_gpu(oracle, cuda_lib):
    _build(oracle, "test_adapter_nrsfm.cc", EXE2)
    rc, out = _run(EXE2)
    assert rc == 0, out
    assert "nrsfm adapter ok" in out


EXE3 = os.path.join(ROOT, "tests", "_emu", "test_adapter_matcher")


def test_matcher_adapter_compiles_and_fails_loudly_without_a_device(oracle, cuda_lib):
    """adapter/MatcherB200.h: DefORBmatcher::SearchByProjection against mock Frame / MapPoint types."""
    _build(oracle, "test_adapter_matcher.cc", EXE3)
    if cuda_lib.defslam_device_count() > 0:
        pytest.skip("device present: covered by the gpu test")
    rc, out = _run(EXE3)
    assert rc == 0, out
    assert "untouched" in out


@pytest.mark.gpu
def test_matcher_adapter_matches_oracle_on_gpu(oracle, cuda_lib):
    _build(oracle, "test_adapter_matcher.cc", EXE3)
    rc, out = _run(EXE3)
    assert rc == 0, out
    assert "0 mismatches" in out


EXE4 = os.path.join(ROOT, "tests", "_emu", "test_adapter_findbywarp")


def test_findbywarp_adapter_compiles_and_fails_loudly_without_a_device(oracle, cuda_lib):
    """adapter/MatcherB200.h: DefORBmatcher::findbyWarp / CalculateInitialSchwarp / searchBySchwarp against mock
    DefKeyFrame / MapPoint types."""
    _build(oracle, "test_adapter_findbywarp.cc", EXE4)
    if cuda_lib.defslam_device_count() > 0:
        pytest.skip("device present: covered by the gpu test")
    rc, out = _run(EXE4)
    assert rc == 0, out
    assert "untouched" in out


@pytest.mark.gpu
def test_findbywarp_adapter_matches_oracle_on_gpu(oracle, cuda_lib):
    _build(oracle, "test_adapter_findbywarp.cc", EXE4)
    rc, out = _run(EXE4)
    assert rc == 0, out
    assert "identical" in out


EXE5 = os.path.join(ROOT, "tests", "_emu", "test_adapter_template")


def test_template_adapter_compiles_and_fails_loudly_without_a_device(oracle, cuda_lib):
    """adapter/TemplateB200.h: TemplateGenerator::LaplacianMeshCreate (LaplacianMesh / TriangularMesh constructors)
    against mock Surface / KeyFrame / Template / Node / Facet / DefMapPoint types."""
    _build(oracle, "test_adapter_template.cc", EXE5)
    if cuda_lib.defslam_device_count() > 0:
        pytest.skip("device present: covered by the gpu test")
    rc, out = _run(EXE5)
    assert rc == 0, out
    assert "untouched" in out


@pytest.mark.gpu
def test_template_adapter_matches_oracle_on_gpu(oracle, cuda_lib):
    _build(oracle, "test_adapter_template.cc", EXE5)
    rc, out = _run(EXE5)
    assert rc == 0, out
    assert "map point diff 0" in out
