"""CPU: the reference's own unit tests for the modules on this path (tests/reference_compat_cases.py) over
tests/fake_backend.py -- exercises the mirror's constructors, call signatures and shapes without a GPU."""
import pytest

from tests import reference_compat_cases as R
from tests.fake_backend import FakeBackend


@pytest.fixture(autouse=True)
def fake_backend(monkeypatch):
    from neurad_studio_b200 import nerfstudio_api

    be = FakeBackend()
    monkeypatch.setattr(nerfstudio_api, "get_backend", lambda device: be)
    return be


@pytest.mark.parametrize("kind", sorted(R.SPACED))
def test_spaced_sampler(kind):
    R.spaced_sampler(R.SPACED[kind], "cpu")


def test_pdf_sampler():
    R.pdf_sampler("cpu")


def test_renderers():
    R.rgb_renderer("cpu")
    R.acc_renderer("cpu")


def test_frustum_get_position():
    R.frustum_get_position("cpu")


def test_spherical_harmonics():
    R.spherical_harmonics("cpu")


def test_tensor_hash_encoder():
    R.tensor_hash_encoder("cpu")


def test_mlp():
    R.mlp("cpu")


def test_standalone_modules_train():
    R.standalone_modules_train("cpu")
